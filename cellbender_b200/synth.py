"""Synthetic 10x-style datasets of BASELINE.json's shapes, and the prepared-dataset object the hot path consumes.

The generator is a chunked, device-side restatement of the reference's simulation recipe
(cellbender/remove_background/data/extras/simulate.py: generate_sample_dirichlet_dataset :217-371,
sample_from_dirichlet_model :716-816): per cell type chi_t ~ Dirichlet(0.05); per cell chi ~ Dirichlet(chi_t * 5000),
epsilon ~ Gamma(100, 100), d ~ LogNormal(log 5000, 0.2), v ~ LogNormal(log 200, 0.4), rho ~ Beta(3, 80), real counts
~ Poisson(mu), background ~ Poisson(lambda) with calculate_mu/lambda('full'), chi_ambient = chi_bar = UMI-weighted mean
of the cell-type profiles.  The reference allocates two dense [n_droplets, n_genes] float64 arrays (43 GB at the PBMC8k
shape); here cells are sampled in chunks and empty droplets through the equivalent Poisson-total + multinomial route.

``prepare_dataset`` performs the part of SingleCellRNACountsDataset that defines the hot path's inputs when
--expected-cells and --total-droplets-included are given (data/dataset.py:115-178, 314-423; data/priors.py:132-175,
364-391): barcode ranking/trimming, count priors, crossover, chi_ambient / chi_bar.  Feature trimming is skipped on
purpose: the benchmark shape is the full 33,538 genes.  Dataset preparation is host-side, one-time and outside the
accelerated path.
"""
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

import numpy as np
import scipy.sparse as sp
import torch


@dataclass
class PreparedDataset:
    """What run_inference reads from SingleCellRNACountsDataset (run.py:696-758)."""

    matrix: sp.csr_matrix  # all simulated droplets x all genes
    analyzed_barcode_inds: np.ndarray
    empty_barcode_inds: np.ndarray
    analyzed_gene_inds: np.ndarray
    priors: Dict[str, Any]
    empty_UMI_threshold: float
    truth: Dict[str, Any] = field(default_factory=dict)

    def get_count_matrix(self) -> sp.csr_matrix:
        return self.matrix[self.analyzed_barcode_inds, :][:, self.analyzed_gene_inds].tocsr()

    def get_count_matrix_empties(self) -> sp.csr_matrix:
        return self.matrix[self.empty_barcode_inds, :][:, self.analyzed_gene_inds].tocsr()

    def restore_eliminated_features_in_cells(self, inferred_count_matrix, cell_probabilities_analyzed_bcs) -> sp.csc_matrix:
        """reference data/dataset.py:495-528 (device pass, sparse_utils.restore_eliminated_features_in_cells)."""
        from .sparse_utils import restore_eliminated_features_in_cells

        return restore_eliminated_features_in_cells(inferred_count_matrix, self.matrix, self.analyzed_gene_inds,
                                                    self.analyzed_barcode_inds, cell_probabilities_analyzed_bcs)


def _coo_rows_from_draws(row_of_draw: torch.Tensor, gene_of_draw: torch.Tensor, n_rows: int, n_genes: int):
    """(row, gene) multiset -> CSR pieces via one sort + unique_consecutive."""
    key = row_of_draw.long() * n_genes + gene_of_draw.long()
    key, _ = torch.sort(key)
    uniq, cnt = torch.unique_consecutive(key, return_counts=True)
    rows = (uniq // n_genes).cpu().numpy()
    cols = (uniq % n_genes).cpu().numpy().astype(np.int32)
    indptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.add.at(indptr, rows + 1, 1)
    return np.cumsum(indptr), cols, cnt.cpu().numpy().astype(np.float32)


@torch.no_grad()
def simulate_dataset(n_genes: int = 1000, cells_of_each_type: List[int] = (334, 333, 333), n_droplets: int = 10000,
                     cell_mean_umi: float = 5000.0, cell_lognormal_sigma: float = 0.2, empty_mean_umi: float = 200.0,
                     empty_lognormal_sigma: float = 0.4, dirichlet_alpha: float = 0.05, alpha0: float = 5000.0,
                     epsilon_param: float = 100.0, rho_alpha: float = 3.0, rho_beta: float = 80.0, seed: int = 0,
                     device: Optional[str] = None, cell_chunk: int = 1000) -> Dict[str, Any]:
    """Dirichlet-Poisson 'full'-model simulation.  Returns dict(matrix CSR [n_droplets, n_genes], chi_ambient, labels ...)."""
    device = device or ("cuda" if torch.cuda.is_available() else "cpu")
    gen = torch.Generator(device=device).manual_seed(seed)
    G = n_genes
    n_types = len(cells_of_each_type)
    n_cells = int(sum(cells_of_each_type))
    assert n_droplets > n_cells

    def gamma(conc):
        return torch._standard_gamma(conc, generator=gen).clamp_min(1e-37)

    chi_types = gamma(torch.full((n_types, G), dirichlet_alpha, device=device, dtype=torch.float64))
    chi_types = chi_types / chi_types.sum(-1, keepdim=True)
    wts = torch.tensor([c * cell_mean_umi for c in cells_of_each_type], device=device, dtype=torch.float64)
    chi_ambient = (chi_types * wts[:, None]).sum(0)
    chi_ambient = chi_ambient / chi_ambient.sum()

    def latents(n):
        eps = gamma(torch.full((n,), epsilon_param, device=device, dtype=torch.float64)) / epsilon_param
        v = torch.exp(np.log(empty_mean_umi) + empty_lognormal_sigma * torch.randn(n, device=device, dtype=torch.float64,
                                                                                     generator=gen))
        a = gamma(torch.full((n,), rho_alpha, device=device, dtype=torch.float64))
        b = gamma(torch.full((n,), rho_beta, device=device, dtype=torch.float64))
        return eps, v, a / (a + b)

    indptrs, cols, vals = [np.zeros(1, dtype=np.int64)], [], []
    labels = np.zeros(n_droplets, dtype=np.int64)
    row0 = 0
    # ---- cells: dense chunks on the device ----
    for t_idx, n_t in enumerate(cells_of_each_type):
        done = 0
        while done < n_t:
            n = min(cell_chunk, n_t - done)
            chi = gamma((chi_types[t_idx] * alpha0 + 1e-20).expand(n, G).float()).double()
            chi = chi / chi.sum(-1, keepdim=True)
            eps, v, rho = latents(n)
            d = torch.exp(np.log(cell_mean_umi) + cell_lognormal_sigma * torch.randn(n, device=device, dtype=torch.float64,
                                                                                       generator=gen))
            mu = ((1 - rho) * eps * d)[:, None] * chi
            lam = eps[:, None] * (((1 - rho) * v)[:, None] * chi_ambient + (rho * (d + v))[:, None] * chi_ambient)
            counts = torch.poisson((mu + lam).float(), generator=gen)
            nz = counts.nonzero(as_tuple=False)
            r = nz[:, 0].cpu().numpy()
            ip = np.zeros(n + 1, dtype=np.int64)
            np.add.at(ip, r + 1, 1)
            indptrs.append(np.cumsum(ip)[1:] + indptrs[-1][-1])
            cols.append(nz[:, 1].cpu().numpy().astype(np.int32))
            vals.append(counts[nz[:, 0], nz[:, 1]].cpu().numpy().astype(np.float32))
            labels[row0:row0 + n] = t_idx + 1
            row0 += n
            done += n
    # ---- empty droplets: Poisson total + multinomial over the ambient profile (== independent Poissons) ----
    n_empty = n_droplets - n_cells
    eps, v, _rho = latents(n_empty)
    totals = torch.poisson((eps * v).float(), generator=gen).long()  # lambda sums to eps * v when chi_bar = chi_ambient
    n_draws = int(totals.sum().item())
    genes = torch.multinomial(chi_ambient.float(), n_draws, replacement=True, generator=gen)
    rows = torch.repeat_interleave(torch.arange(n_empty, device=device), totals)
    ip, c, d_ = _coo_rows_from_draws(rows, genes, n_empty, G)
    indptrs.append(ip[1:] + indptrs[-1][-1])
    cols.append(c)
    vals.append(d_)
    matrix = sp.csr_matrix((np.concatenate(vals), np.concatenate(cols), np.concatenate(indptrs)), shape=(n_droplets, G))
    return {"matrix": matrix, "chi_ambient": chi_ambient.cpu().numpy(), "droplet_labels": labels,
            "chi": chi_types.cpu().numpy(), "n_cells": n_cells}


def simulate_dataset_host(n_genes: int = 1000, cells_of_each_type: List[int] = (334, 333, 333), n_droplets: int = 10000,
                          cell_mean_umi: float = 5000.0, cell_lognormal_sigma: float = 0.2, empty_mean_umi: float = 200.0,
                          empty_lognormal_sigma: float = 0.4, dirichlet_alpha: float = 0.05, alpha0: float = 5000.0,
                          epsilon_param: float = 100.0, rho_alpha: float = 3.0, rho_beta: float = 80.0,
                          seed: int = 0) -> Dict[str, Any]:
    """The same Dirichlet-Poisson 'full'-model recipe, dense on the host with ONE ``np.random.RandomState(seed)`` stream
    (legacy numpy generators are bit-reproducible across machines, torch's device generators are not): the datasets of
    the end-to-end parity fixtures (tests/golden/e2e_*.npz) are regenerated from the seed on the GPU box and must equal
    what the golden run saw in the build container.  Small shapes only (two dense [n_droplets, n_genes] arrays).
    Additionally returns the true / background count matrices."""
    rng = np.random.RandomState(seed)
    G, n_types, n_cells = n_genes, len(cells_of_each_type), int(sum(cells_of_each_type))
    assert n_droplets > n_cells
    chi_types = rng.dirichlet(np.full(G, dirichlet_alpha), size=n_types)
    chi_ambient = (chi_types * (np.asarray(cells_of_each_type, dtype=np.float64) * cell_mean_umi)[:, None]).sum(0)
    chi_ambient = chi_ambient / chi_ambient.sum()
    c_real = np.zeros((n_droplets, G), dtype=np.int64)
    c_bkg = np.zeros((n_droplets, G), dtype=np.int64)
    labels = np.zeros(n_droplets, dtype=np.int64)

    def latents(n):
        eps = rng.gamma(shape=epsilon_param, scale=1.0 / epsilon_param, size=n)
        v = rng.lognormal(mean=np.log(empty_mean_umi), sigma=empty_lognormal_sigma, size=n)
        return eps, v, rng.beta(a=rho_alpha, b=rho_beta, size=n)

    row = 0
    for t, n in enumerate(cells_of_each_type):
        chi = rng.dirichlet(chi_types[t] * alpha0 + 1e-20, size=n)
        eps, v, rho = latents(n)
        d = rng.lognormal(mean=np.log(cell_mean_umi), sigma=cell_lognormal_sigma, size=n)
        c_real[row:row + n] = rng.poisson(((1 - rho) * eps * d)[:, None] * chi)
        c_bkg[row:row + n] = rng.poisson((eps * ((1 - rho) * v + rho * (d + v)))[:, None] * chi_ambient[None, :])
        labels[row:row + n] = t + 1
        row += n
    eps, v, rho = latents(n_droplets - row)
    c_bkg[row:] = rng.poisson((eps * v)[:, None] * chi_ambient[None, :])  # y = 0: lambda = eps * v * chi_ambient
    matrix = sp.csr_matrix((c_real + c_bkg).astype(np.float32))
    return {"matrix": matrix, "chi_ambient": chi_ambient, "droplet_labels": labels, "chi": chi_types, "n_cells": n_cells,
            "counts_true": sp.csr_matrix(c_real), "counts_bkg": sp.csr_matrix(c_bkg)}


def dataset_checksum(matrix: sp.csr_matrix) -> int:
    """Order-sensitive 64-bit checksum of a canonical CSR matrix (fixture guard: same data here and on the GPU box)."""
    m = sp.csr_matrix(matrix)
    m.sum_duplicates()
    m.sort_indices()
    acc = np.uint64(1469598103934665603)
    with np.errstate(over="ignore"):
        for arr in (m.indptr.astype(np.uint64), m.indices.astype(np.uint64), m.data.astype(np.int64).astype(np.uint64)):
            w = (np.arange(arr.size, dtype=np.uint64) * np.uint64(2654435761) + np.uint64(40503))
            acc = acc * np.uint64(1099511628211) + np.uint64((arr * w).sum(dtype=np.uint64))
    return int(acc)


def prepare_dataset(sim: Dict[str, Any], expected_cells: int, total_droplets_included: int,
                    low_count_threshold: int = 5, fraction_empties: float = 0.2,
                    analyzed_barcode_inds: Optional[np.ndarray] = None,
                    empty_barcode_inds: Optional[np.ndarray] = None) -> PreparedDataset:
    """Barcode trimming and priors as the reference computes them when --expected-cells and
    --total-droplets-included are supplied.  ``analyzed_barcode_inds`` / ``empty_barcode_inds`` pin the barcode selection
    (fixtures: the order of equal UMI counts under an unstable argsort is not portable across CPUs)."""
    matrix: sp.csr_matrix = sim["matrix"].tocsr()
    umi = np.asarray(matrix.sum(axis=1)).ravel()
    order = np.argsort(umi)[::-1]
    priors: Dict[str, Any] = {"expected_cells": int(expected_cells), "total_droplets": int(total_droplets_included)}
    # priors.py:132-145, 148-175
    priors["cell_counts"] = float(np.exp(np.mean(np.log(umi[order][:expected_cells]))))
    start = max(expected_cells, total_droplets_included - 500)
    priors["empty_counts"] = float(np.median(umi[order][int(start):int(total_droplets_included)]))
    middle = np.sqrt(priors["cell_counts"] * priors["empty_counts"])
    priors["empty_count_upper_limit"] = float(min(middle, 1.5 * priors["empty_counts"]))
    # priors.py:364-391
    rs = np.sort(umi)[::-1]
    surely_empty_counts = rs[total_droplets_included]
    priors["surely_empty_counts"] = surely_empty_counts
    priors["log_counts_crossover"] = float((np.log(surely_empty_counts) + np.log(priors["cell_counts"])) / 2)
    log_nz = np.log(umi[umi > 0])
    priors["d_std"] = float(np.std(log_nz[log_nz > priors["log_counts_crossover"]]) / 5.0)
    priors["d_empty_std"] = 0.01
    # dataset.py:314-404
    low_count_cutoff = max(low_count_threshold, int(priors["empty_counts"] * 0.5))
    n_above = int(np.sum(umi > low_count_cutoff))
    assert n_above > expected_cells, "no empty droplets above the low-count cutoff"
    analyzed = order[:total_droplets_included]
    empties = order[np.arange(total_droplets_included, n_above, dtype=int)]
    empty_UMI_threshold = float(umi[order][total_droplets_included])
    cell_prob = (expected_cells / total_droplets_included) * (1.0 - fraction_empties)
    priors["cell_logit"] = float(np.log(cell_prob) - np.log(1.0 - cell_prob))
    # dataset.py:406-423
    ep = np.finfo(np.float32).eps.item()
    empty_logic = (umi < priors["surely_empty_counts"]) & (umi > low_count_cutoff)
    amb = np.asarray(matrix[empty_logic, :].sum(axis=0)).ravel() + ep
    priors["chi_ambient"] = torch.tensor(amb / amb.sum()).float()
    bar = np.asarray(matrix.sum(axis=0)).ravel() + ep
    priors["chi_bar"] = torch.tensor(bar / bar.sum()).float()
    if analyzed_barcode_inds is not None:
        analyzed, empties = np.asarray(analyzed_barcode_inds, dtype=np.int64), np.asarray(empty_barcode_inds, dtype=np.int64)
    return PreparedDataset(matrix=matrix, analyzed_barcode_inds=np.asarray(analyzed), empty_barcode_inds=np.asarray(empties),
                           analyzed_gene_inds=np.arange(matrix.shape[1]), priors=priors,
                           empty_UMI_threshold=empty_UMI_threshold,
                           truth={k: sim[k] for k in ("chi_ambient", "droplet_labels", "n_cells", "counts_true", "counts_bkg",
                                                        "bkg_per_droplet") if k in sim})


def save_fixture_dataset(path: str, ds: PreparedDataset, expected_cells: int, total_droplets: int):
    """The raw count matrix + the barcode selection of a prepared (small) dataset as a compressed .npz fixture."""
    m = sp.csr_matrix(ds.matrix)
    m.sort_indices()
    assert m.data.max() < 65536 and m.shape[1] < 32768
    bkg = np.asarray(ds.truth["counts_bkg"].sum(axis=1)).ravel() if "counts_bkg" in ds.truth else np.zeros(m.shape[0])
    np.savez_compressed(path, indptr=m.indptr.astype(np.int32), indices=m.indices.astype(np.int16), data=m.data.astype(np.uint16),
                        shape=np.array(m.shape), analyzed_barcode_inds=np.asarray(ds.analyzed_barcode_inds).astype(np.int32),
                        empty_barcode_inds=np.asarray(ds.empty_barcode_inds).astype(np.int32), bkg_per_droplet=bkg.astype(np.int32),
                        expected_cells=np.array(expected_cells), total_droplets=np.array(total_droplets))


def load_fixture_dataset(path: str) -> PreparedDataset:
    """Inverse of save_fixture_dataset: priors are recomputed from the stored matrix (deterministic), the barcode
    selection is the stored one."""
    f = np.load(path)
    matrix = sp.csr_matrix((f["data"].astype(np.float32), f["indices"].astype(np.int32), f["indptr"].astype(np.int64)),
                           shape=tuple(int(v) for v in f["shape"]))
    sim = {"matrix": matrix, "chi_ambient": None, "droplet_labels": None, "n_cells": int(f["expected_cells"]),
           "bkg_per_droplet": f["bkg_per_droplet"].astype(np.int64)}
    return prepare_dataset(sim, int(f["expected_cells"]), int(f["total_droplets"]),
                           analyzed_barcode_inds=f["analyzed_barcode_inds"], empty_barcode_inds=f["empty_barcode_inds"])


CONFIGS = {
    # BASELINE.json configs[0..3]; "tiny" is the reference's own test fixture shape (tests/conftest.py:77-105)
    "tiny": dict(n_genes=1000, cells_of_each_type=[10, 10, 10], n_droplets=100 + 400, expected_cells=30, total_droplets=100),
    "c1": dict(n_genes=1000, cells_of_each_type=[334, 333, 333], n_droplets=10000, expected_cells=1000, total_droplets=3000),
    "pbmc8k": dict(n_genes=33538, cells_of_each_type=[2000] * 4, n_droplets=75000, expected_cells=8000, total_droplets=25000),
    "citeseq": dict(n_genes=36200, cells_of_each_type=[2500] * 4, n_droplets=80000, expected_cells=10000, total_droplets=30000),
    "atlas": dict(n_genes=36601, cells_of_each_type=[12500] * 8, n_droplets=400000, expected_cells=100000, total_droplets=300000),
}


def make_config_dataset(name: str, seed: int = 0, device: Optional[str] = None, host: bool = False) -> PreparedDataset:
    """``host=True``: machine-independent numpy generation (small shapes; the end-to-end parity fixtures)."""
    cfg = dict(CONFIGS[name])
    expected_cells, total = cfg.pop("expected_cells"), cfg.pop("total_droplets")
    sim = simulate_dataset_host(seed=seed, **cfg) if host else simulate_dataset(seed=seed, device=device, **cfg)
    return prepare_dataset(sim, expected_cells, total)
