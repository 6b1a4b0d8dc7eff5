"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference
(imported from /root/reference through oracle/ref_shim.py) on seeded inputs.

Run in the build container only (the reference does not travel to the GPU box):
    python tools/make_golden.py
The outputs are small .npz / .pt fixtures that are committed.  tests/test_oracle_golden.py
pins the oracle against them; the -m gpu tests pin the CUDA path against the oracle and,
for the estimators, directly against these reference outputs.
"""
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

ref_shim.install()

from cellbender.remove_background import consts  # noqa: E402
from cellbender.remove_background.data.dataprep import DataLoader  # noqa: E402
from cellbender.remove_background.distributions.NegativeBinomialPoissonConv import (  # noqa: E402
    NegativeBinomialPoissonConv as NBPC,
)
from cellbender.remove_background.distributions.NegativeBinomialPoissonConvApprox import (  # noqa: E402
    NegativeBinomialPoissonConvApprox as NBPCapprox,
)
from cellbender.remove_background.estimation import MAP, Mean, MultipleChoiceKnapsack, ThresholdCDF  # noqa: E402
from cellbender.remove_background.model import calculate_lambda, calculate_mu, get_p_logit_prior  # noqa: E402
from cellbender.remove_background.posterior import (  # noqa: E402
    IndexConverter,
    Posterior,
    PRmu,
    PRq,
    compute_mean_target_removal_as_function,
    torch_binary_search,
)
from cellbender.remove_background.sparse_utils import dense_to_sparse_op_torch  # noqa: E402
from cellbender.remove_background.vae.decoder import Decoder  # noqa: E402
from cellbender.remove_background.vae.encoder import EncodeNonZLatents, EncodeZ  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def golden_nbpc():
    g = torch.Generator().manual_seed(11)
    n = 4096
    # wide dynamic range: mu from 1e-10 (the y=0 safeguard) to 1e3, lam 1e-8..1e2, counts 0..600
    mu = torch.exp(torch.empty(n, dtype=torch.float64).uniform_(np.log(1e-10), np.log(1e3), generator=g))
    mu[:256] = 1e-10
    lam = torch.exp(torch.empty(n, dtype=torch.float64).uniform_(np.log(1e-8), np.log(1e2), generator=g))
    alpha = torch.exp(torch.empty(n, dtype=torch.float64).uniform_(np.log(0.5), np.log(50.0), generator=g))
    value = torch.poisson((mu + lam).clamp(max=500.0), generator=g)
    value[torch.rand(n, generator=g) < 0.5] = 0.0
    value[-8:] = torch.tensor([1, 2, 3, 50, 100, 200, 400, 600], dtype=torch.float64)
    # inputs are made exactly fp32-representable so fp32 / fp64 / multiprecision all see the same numbers
    mu, lam, alpha, value = (t.float().double() for t in (mu, lam, alpha, value))
    out = {}
    out.update(_nbpc_multiprecision(mu.numpy(), alpha.numpy(), lam.numpy(), value.numpy()))
    for name, dt in (("f64", torch.float64), ("f32", torch.float32)):
        m_, a_, l_, v_ = (t.to(dt).clone() for t in (mu, alpha, lam, value))
        m_.requires_grad_(True), a_.requires_grad_(True), l_.requires_grad_(True)
        lp = NBPCapprox(mu=m_, alpha=a_, lam=l_).log_prob(v_)
        gm, ga, gl = torch.autograd.grad(lp.sum(), (m_, a_, l_))
        out[f"approx_{name}"] = lp.detach().numpy()
        out[f"approx_dmu_{name}"] = gm.numpy()
        out[f"approx_dalpha_{name}"] = ga.numpy()
        out[f"approx_dlam_{name}"] = gl.numpy()
    # exact convolution (positive mu only, moderate counts: the reference table is [n, 101])
    ok = (mu > 1e-6)
    lpx = NBPC(mu=mu[ok], alpha=alpha[ok], lam=lam[ok], max_poisson=100).log_prob(value[ok])
    out["exact_mask"] = ok.numpy()
    out["exact_f64"] = lpx.numpy()
    np.savez_compressed(os.path.join(OUT, "nbpc.npz"), mu=mu.numpy(), alpha=alpha.numpy(), lam=lam.numpy(),
                        value=value.numpy(), **out)


def _nbpc_multiprecision(mu, alpha, lam, value):
    """Multiprecision (mpmath) evaluation of the reference formula ...ConvApprox.py:63-70,117-127 and of its
    analytic derivatives; see oracle/nbpc.py:nbpc_approx_mp."""
    from oracle.nbpc import nbpc_approx_mp

    L, dmu, dlam, dal = nbpc_approx_mp(value, mu, alpha, lam, dps=80)
    return {"approx_mp": L, "approx_dmu_mp": dmu, "approx_dlam_mp": dlam, "approx_dalpha_mp": dal}


def golden_rates():
    g = torch.Generator().manual_seed(5)
    B, G = 7, 33
    eps = torch.rand(B, generator=g, dtype=torch.float64) + 0.5
    d_cell = torch.rand(B, generator=g, dtype=torch.float64) * 5000 + 100
    d_empty = torch.rand(B, generator=g, dtype=torch.float64) * 200 + 20
    rho = torch.rand(B, generator=g, dtype=torch.float64) * 0.1
    chi = torch.softmax(torch.randn(B, G, generator=g, dtype=torch.float64) * 3, -1)
    chi_amb = torch.softmax(torch.randn(G, generator=g, dtype=torch.float64), -1)
    chi_bar = torch.softmax(torch.randn(G, generator=g, dtype=torch.float64), -1)
    y = (torch.rand(B, generator=g) < 0.5).double()
    out = dict(eps=eps, d_cell=d_cell, d_empty=d_empty, rho=rho, chi=chi, chi_amb=chi_amb, chi_bar=chi_bar, y=y)
    for mt in ("simple", "ambient", "swapping", "full"):
        out[f"mu_{mt}"] = calculate_mu(mt, epsilon=eps, d_cell=d_cell, chi=chi, y=y, rho=rho)
        out[f"lam_{mt}"] = calculate_lambda(mt, epsilon=eps, chi_ambient=chi_amb, d_empty=d_empty, y=y,
                                            d_cell=d_cell, rho=rho, chi_bar=chi_bar)
    lc = torch.tensor([2.0, 4.0, 5.5, 6.0, 7.5, 8.5, 9.2], dtype=torch.float64)
    out["plp_log_counts"] = lc
    out["plp"] = get_p_logit_prior(lc, torch.tensor(8.5, dtype=torch.float64), 50, -1.3)
    np.savez_compressed(os.path.join(OUT, "rates.npz"), **{k: v.numpy() for k, v in out.items()})


def golden_posterior():
    g = torch.Generator().manual_seed(3)
    res = {}
    for tag, (B, G, scale) in {"big": (6, 48, 30.0), "small": (4, 24, 2.0), "huge": (3, 16, 800.0)}.items():
        mu = torch.exp(torch.randn(B, G, generator=g, dtype=torch.float64)) * scale
        mu[0, :4] = 0.0  # exercises the +1e-30 safeguard path (y = 0 droplets)
        lam = torch.exp(torch.randn(B, G, generator=g, dtype=torch.float64)) * scale * 0.1
        alpha = torch.tensor(4.7, dtype=torch.float64)
        data = torch.poisson(mu + lam, generator=g)
        data[torch.rand(B, G, generator=g) < 0.4] = 0.0
        for name, dt in (("f64", torch.float64), ("f32", torch.float32)):
            lp, low = Posterior._log_prob_noise_count_tensor(
                data=data.to(dt), mu_est=(mu + 1e-30).to(dt), lambda_est=(lam + 1e-30).to(dt),
                alpha_est=(alpha + 1e-30).to(dt), n_counts_max=20)
            res[f"{tag}_lp_{name}"] = lp.numpy()
            res[f"{tag}_low_{name}"] = low.numpy()
        res[f"{tag}_mu"], res[f"{tag}_lam"], res[f"{tag}_data"] = mu.numpy(), lam.numpy(), data.numpy()
        res[f"{tag}_alpha"] = alpha.numpy()
    np.savez_compressed(os.path.join(OUT, "posterior_logprob.npz"), **res)


def _fixture_posterior():
    """The known-answer posterior of the reference's tests/test_estimation.py:25-59 (re-derived)."""
    n = -np.inf
    rest = np.log(1.0 - np.exp([-0.74, -1, -2, -4]).sum())
    m = np.array([
        [0, n, n, n, n, n, n, n, n, n],
        [n, 0, n, n, n, n, n, n, n, n],
        [n, n, 0, n, n, n, n, n, n, n],
        [-6, -2, np.log(1.0 - 2 * np.exp(-2) - np.exp(-6)), -2, n, n, n, n, n, n],
        [-2.5, -1.5, -0.5, -3, np.log(1.0 - np.exp([-2.5, -1.5, -0.5, -3]).sum()), n, n, n, n, n],
        [-0.74, -1, -2, -4, rest, n, n, n, n, n],
        [-1, -0.74, -2, -4, rest, n, n, n, n, n],
        [-2, -1, -0.74, -4, rest, n, n, n, n, n],
    ])
    rows, cols, vals = dense_to_sparse_op_torch(torch.tensor(m), tensor_for_nonzeros=torch.tensor(m).exp())
    coo = sp.coo_matrix((vals.numpy(), (rows.numpy(), cols.numpy())), shape=m.shape)
    offsets = dict(zip(range(8), [0] * 7 + [1]))
    return coo, offsets


def _random_posterior(rng, n_cells, n_genes, density=0.5, n_c=12, with_holes=True):
    """Random posterior COO: normalised log probs over a window, thresholded, shuffled."""
    rows, cols, vals, offsets = [], [], [], {}
    for m in range(n_cells * n_genes):
        if rng.random() > density:
            continue
        k = rng.integers(1, n_c + 1)
        lp = rng.normal(size=k) * 2.0
        if rng.random() < 0.3:
            lp = np.round(lp)  # ties in argmax and in delta
        lp = lp - np.log(np.exp(lp).sum())
        keep = lp >= -6 if with_holes else np.ones(k, bool)
        keep[np.argmax(lp)] = True
        for c in np.flatnonzero(keep):
            rows.append(m), cols.append(c), vals.append(lp[c])
        if rng.random() < 0.3:
            offsets[m] = int(rng.integers(1, 4))
    rows, cols, vals = np.array(rows), np.array(cols), np.array(vals, dtype=np.float32)
    # make sure every column value occurs (the reference's compact-column quirk)
    perm = rng.permutation(rows.size)
    coo = sp.coo_matrix((vals[perm], (rows[perm], cols[perm])), shape=(n_cells * n_genes, n_c))
    return coo, offsets


def golden_estimation():
    res = {}
    coo, offsets = _fixture_posterior()
    cases = {"fixture": (coo, offsets, 1, 8)}
    rng = np.random.default_rng(7)
    for i in range(6):
        nc, ng = int(rng.integers(3, 9)), int(rng.integers(4, 12))
        c, o = _random_posterior(rng, nc, ng)
        cases[f"rand{i}"] = (c, o, nc, ng)
    for tag, (c, o, nc, ng) in cases.items():
        conv = IndexConverter(total_n_cells=nc, total_n_genes=ng)
        res[f"{tag}_row"], res[f"{tag}_col"], res[f"{tag}_data"] = c.row.astype(np.int64), c.col.astype(np.int32), c.data
        res[f"{tag}_shape"] = np.array([nc, ng, c.shape[1]])
        res[f"{tag}_off_m"] = np.array(list(o.keys()), dtype=np.int64)
        res[f"{tag}_off_v"] = np.array(list(o.values()), dtype=np.int64)
        res[f"{tag}_map"] = np.array(MAP(conv).estimate_noise(c, o).todense())
        res[f"{tag}_mean"] = np.array(Mean(conv).estimate_noise(c, o).todense())
        for q in (0.5, 0.9):
            res[f"{tag}_cdf{int(q * 100)}"] = np.array(ThresholdCDF(conv).estimate_noise(c, o, q=q).todense())
        map_per_gene = res[f"{tag}_map"].sum(axis=0)
        tg = {"t0": np.zeros(ng), "tmap": map_per_gene.astype(float), "tplus": map_per_gene + 2.6,
              "tminus": map_per_gene - 1.5, "tmix": map_per_gene + rng.integers(-4, 5, size=ng) + 0.5,
              "tbig": map_per_gene + 1000.0}
        for tname, t in tg.items():
            res[f"{tag}_mckp_{tname}_target"] = t
            res[f"{tag}_mckp_{tname}"] = np.array(
                MultipleChoiceKnapsack(conv).estimate_noise(c, o, noise_targets_per_gene=t, n_chunks=1).todense())
        # mean-based per-gene target function (posterior.py:1589-1650)
        raw = sp.csr_matrix(rng.poisson(3.0, size=(nc, ng)).astype(np.int32))
        fn = compute_mean_target_removal_as_function(c, o, conv, raw, n_cells=nc, device="cpu", per_gene=True)
        res[f"{tag}_raw"] = np.array(raw.todense())
        res[f"{tag}_target_fpr01"] = fn(0.01).numpy()
    # the reference's own MCKP truth tables (tests/test_estimation.py:289-305) on the fixture
    np.savez_compressed(os.path.join(OUT, "estimation.npz"), **res)


def golden_prq():
    """PRq.regularize (posterior.py:1002-1225) of the unmodified reference on its own test fixture and on random
    posteriors with holes and offsets, alpha in {0, 1, 2}; plus torch_binary_search on the reference test's function."""
    res = {}
    coo, offsets = _fixture_posterior()
    cases = {"fixture": (coo, offsets)}
    rng = np.random.default_rng(11)
    for i in range(3):
        nc, ng = int(rng.integers(3, 9)), int(rng.integers(4, 12))
        cases[f"rand{i}"] = _random_posterior(rng, nc, ng)
    for tag, (coo, offsets) in cases.items():
        res[f"{tag}_row"], res[f"{tag}_col"], res[f"{tag}_data"] = coo.row.astype(np.int64), coo.col.astype(np.int32), coo.data
        res[f"{tag}_shape"] = np.array(coo.shape)
        off_m = np.array(sorted(offsets.keys()), dtype=np.int64)
        res[f"{tag}_off_m"], res[f"{tag}_off_v"] = off_m, np.array([offsets[m] for m in off_m], dtype=np.int32)
        for alpha in (0.0, 1.0, 2.0):
            tgt = PRq._compute_log_target_dict(noise_count_posterior_coo=coo, alpha=alpha)
            res[f"{tag}_a{alpha:g}_target_m"] = np.array(list(tgt.keys()), dtype=np.int64)
            res[f"{tag}_a{alpha:g}_target"] = np.array(list(tgt.values()), dtype=np.float64)
            reg = PRq.regularize(noise_count_posterior_coo=coo, noise_offsets=offsets, alpha=alpha, device="cpu",
                                 target_tolerance=0.001)
            order = np.lexsort((reg.col, reg.row))
            res[f"{tag}_a{alpha:g}_row"] = reg.row[order].astype(np.int64)
            res[f"{tag}_a{alpha:g}_col"] = reg.col[order].astype(np.int32)
            res[f"{tag}_a{alpha:g}_data"] = np.asarray(reg.data, dtype=np.float64)[order]
    # torch_binary_search (tests/test_posterior.py:280-330 style): monotone function, per-element targets
    x0 = torch.tensor([0.3, 2.0, 7.5, 40.0])
    out = torch_binary_search(evaluate_outcome_given_value=lambda x: x ** 2, target_outcome=x0 ** 2,
                              init_range=torch.tensor([0.0, 100.0]).unsqueeze(0).expand(4, 2), target_tolerance=0.001,
                              max_iterations=100)
    res["bsearch_target"], res["bsearch_out"] = (x0 ** 2).numpy(), out.numpy()
    np.savez_compressed(os.path.join(OUT, "prq.npz"), **res)


def golden_prmu():
    """PRmu.regularize (posterior.py:1228-1530) of the unmodified reference: overall and per-gene mean targeting on
    random posteriors with offsets; n_cells >= the number of cells, so the random subset is every cell."""
    res = {}
    rng = np.random.default_rng(23)
    for i in range(2):
        nc, ng = int(rng.integers(5, 9)), int(rng.integers(6, 12))
        coo, offsets = _random_posterior(rng, nc, ng, density=0.7)
        raw = sp.csr_matrix(rng.poisson(6.0, size=(nc, ng)).astype(np.int64))
        conv = IndexConverter(total_n_cells=nc, total_n_genes=ng)
        tag = f"case{i}"
        res[f"{tag}_row"], res[f"{tag}_col"], res[f"{tag}_data"] = coo.row.astype(np.int64), coo.col.astype(np.int32), coo.data
        res[f"{tag}_shape"], res[f"{tag}_ncng"] = np.array(coo.shape), np.array([nc, ng])
        off_m = np.array(sorted(offsets.keys()), dtype=np.int64)
        res[f"{tag}_off_m"], res[f"{tag}_off_v"] = off_m, np.array([offsets[m] for m in off_m], dtype=np.int32)
        res[f"{tag}_raw"] = np.asarray(raw.todense())
        for per_gene in (False, True):
            for fpr in (0.01, 0.3):
                np.random.seed(0)
                reg = PRmu.regularize(noise_count_posterior_coo=coo, noise_offsets=offsets, index_converter=conv,
                                      raw_count_matrix=raw.copy(), fpr=fpr, per_gene=per_gene, device="cpu", n_cells=1000)
                order = np.lexsort((reg.col, reg.row))
                key = f"{tag}_{'gene' if per_gene else 'all'}_f{fpr:g}"
                res[key + "_row"] = reg.row[order].astype(np.int64)
                res[key + "_col"] = reg.col[order].astype(np.int32)
                res[key + "_data"] = np.asarray(reg.data, dtype=np.float64)[order]
    np.savez_compressed(os.path.join(OUT, "prmu.npz"), **res)


def golden_sparse_utils():
    """Output assembly of the unmodified reference: overwrite_matrix_with_columns_from_another and csr_set_rows_to_zero
    (sparse_utils.py:58-100) on the cases of tests/test_sparse_utils.py:128-139,170-180 and on larger random count
    matrices; `cell_counts - noise_csr` (posterior.py:315-328) and restore_eliminated_features_in_cells
    (data/dataset.py:495-528, called unbound on a namespace carrying the three attributes it reads)."""
    from types import SimpleNamespace

    from cellbender.remove_background.data.dataset import SingleCellRNACountsDataset
    from cellbender.remove_background.sparse_utils import csr_set_rows_to_zero, overwrite_matrix_with_columns_from_another

    rng = np.random.RandomState(7)
    res = {}
    cases = {
        "t0": (np.array([[1, 2], [3, 4]]), np.zeros((2, 2), int), [0]),
        "t1": (np.array([[1, 2], [3, 4], [5, 6]]), np.zeros((3, 2), int), [1]),
        "t2": (rng.poisson(2.0, size=(10, 10)), rng.poisson(2.0, size=(10, 10)), [0, 1, 2, 3]),
        "r0": (rng.poisson(0.3, size=(70, 90)), rng.poisson(0.4, size=(70, 90)), sorted(rng.choice(90, 55, replace=False))),
        "r1": (rng.poisson(0.05, size=(33, 300)), rng.poisson(1.5, size=(33, 300)), sorted(rng.choice(300, 3, replace=False))),
        "none": (rng.poisson(0.5, size=(9, 12)), rng.poisson(0.5, size=(9, 12)), []),
    }
    for tag, (m1, m2, cols) in cases.items():
        out = overwrite_matrix_with_columns_from_another(sp.csc_matrix(m1), sp.csc_matrix(m2), np.array(cols, dtype=int))
        res[f"ow_{tag}_m1"], res[f"ow_{tag}_m2"], res[f"ow_{tag}_cols"] = m1, m2, np.array(cols, dtype=np.int64)
        res[f"ow_{tag}_out"] = np.asarray(out.todense())
        assert sp.issparse(out) and out.format == "csc"
    for tag, (m, rows) in {"t0": (np.array([[1, 2], [3, 4]]), [0]), "t1": (np.array([[1, 2], [3, 4], [5, 6]]), [1]),
                           "r": (rng.poisson(0.4, size=(50, 40)), sorted(rng.choice(50, 21, replace=False)))}.items():
        out = csr_set_rows_to_zero(sp.csr_matrix(m), rows)
        res[f"rz_{tag}_m"], res[f"rz_{tag}_rows"], res[f"rz_{tag}_out"] = m, np.array(rows, dtype=np.int64), np.asarray(out.todense())
        res[f"rz_{tag}_nnz"] = out.nnz
    # denoise + restore: 60 barcodes x 80 features, 45 analysed barcodes, 50 analysed features
    raw = rng.poisson(0.6, size=(60, 80))
    bcs = np.sort(rng.choice(60, 45, replace=False))
    genes = np.sort(rng.choice(80, 50, replace=False))
    p = rng.uniform(size=45)
    p[:3] = [0.5, 0.5000001, 0.4999999]
    cell_inds = bcs[p > consts.CELL_PROB_CUTOFF]
    noise = np.zeros_like(raw)
    sub = np.ix_(cell_inds, genes)
    noise[sub] = rng.binomial(raw[sub], 0.4)  # noise <= count, only where the posterior exists
    count_matrix = sp.csr_matrix(raw)
    empty_inds = set(range(60)) - set(cell_inds)
    denoised = (csr_set_rows_to_zero(csr=count_matrix.copy(), row_inds=empty_inds) - sp.csr_matrix(noise)).tocsc()
    ns = SimpleNamespace(data={"matrix": sp.csc_matrix(raw)}, analyzed_gene_inds=genes, analyzed_barcode_inds=bcs)
    restored = SingleCellRNACountsDataset.restore_eliminated_features_in_cells(ns, denoised, p)
    res.update(dn_raw=raw, dn_bcs=bcs, dn_genes=genes, dn_p=p, dn_noise=noise, dn_denoised=np.asarray(denoised.todense()),
               dn_denoised_nnz=denoised.nnz, dn_restored=np.asarray(restored.todense()), dn_restored_nnz=restored.nnz)
    np.savez_compressed(os.path.join(OUT, "sparse_utils.npz"), **res)


def golden_vae():
    torch.manual_seed(21)
    G, H, Z, B = 40, 16, 6, 12
    import cellbender.remove_background.vae.encoder as encmod

    # the reference hard-codes 512 hidden units in EncodeNonZLatents; keep them.
    enc_z = EncodeZ(input_dim=G, hidden_dims=[H], output_dim=Z, use_batch_norm=False, use_layer_norm=False,
                    input_transform="normalize")
    enc_o = EncodeNonZLatents(n_genes=G, z_dim=Z, log_count_crossover=5.0, prior_log_cell_counts=7.0,
                              empty_log_count_threshold=3.5, prior_logit_cell_prob=-1.0,
                              input_transform="log_normalize")
    dec = Decoder(input_dim=Z, hidden_dims=[H], use_batch_norm=True, use_layer_norm=False, output_dim=G)
    # non-trivial BN state
    for mod in list(enc_o.modules()) + list(dec.modules()):
        if isinstance(mod, torch.nn.BatchNorm1d):
            mod.running_mean.normal_(0, 0.1)
            mod.running_var.uniform_(0.5, 1.5)
            mod.weight.data.uniform_(0.5, 1.5)
            mod.bias.data.normal_(0, 0.1)
    x = torch.poisson(torch.rand(B, G) * 3).float()
    x[:7] *= 40  # cells: larger counts so that > 4 rows exceed the crossover
    chi_amb = torch.softmax(torch.randn(G), -1)
    ref_shim.set_param("d_empty_loc", torch.tensor(4.0))
    enc_o.dropout50 = torch.nn.Identity()  # dropout realisation is an explicit input in the restatement
    out = {"x": x, "chi_amb": chi_amb}
    for mode in ("train", "eval"):
        for mod in (enc_z, enc_o, dec):
            mod.train(mode == "train")
        enc_o.offset = None
        zz = enc_z(x=x)
        oo = enc_o(x=x, chi_ambient=chi_amb, z=zz["loc"].detach())
        out[f"{mode}_z_loc"], out[f"{mode}_z_scale"] = zz["loc"].detach(), zz["scale"].detach()
        for k in ("p_y", "d_loc", "epsilon"):
            out[f"{mode}_{k}"] = oo[k].detach()
        out[f"{mode}_offset_logit_p"] = torch.tensor(enc_o.offset["logit_p"])
        out[f"{mode}_offset_epsilon"] = torch.tensor(enc_o.offset["epsilon"])
        zin = torch.randn(B, Z, generator=torch.Generator().manual_seed(1))
        out[f"{mode}_dec_in"] = zin
        out[f"{mode}_chi"] = dec(zin).detach()
    # NOTE: train-mode forward updates BN running stats; save the state AFTER, and eval
    # outputs above were produced with exactly this state.
    torch.save({"enc_z": enc_z.state_dict(), "enc_o": enc_o.state_dict(), "dec": dec.state_dict(),
                "io": out, "dims": dict(G=G, H=H, Z=Z, B=B),
                "cfg": dict(log_count_crossover=5.0, prior_log_cell_counts=7.0, empty_log_count_threshold=3.5,
                            d_empty_loc=4.0)},
               os.path.join(OUT, "vae.pt"))


def golden_dataloader():
    rng = np.random.default_rng(0)
    dataset = sp.csr_matrix(rng.poisson(0.3, size=(53, 30)).astype(np.int32))
    empties = sp.csr_matrix(rng.poisson(0.05, size=(17, 30)).astype(np.int32))
    np.random.seed(1234)
    loader = DataLoader(dataset, empties, batch_size=16, fraction_empties=0.2, shuffle=True, use_cuda=False)
    n_batches = len(loader)  # consumes RNG exactly as run.py:762 does
    batches = [b.numpy() for b in loader]
    batches2 = [b.numpy() for b in loader]  # second epoch (reshuffled)
    np.savez_compressed(
        os.path.join(OUT, "dataloader.npz"), dataset=np.array(dataset.todense()), empties=np.array(empties.todense()),
        n_batches=n_batches, **{f"e0_b{i}": b for i, b in enumerate(batches)},
        **{f"e1_b{i}": b for i, b in enumerate(batches2)})


if __name__ == "__main__":
    if "--only-sparse" in sys.argv:
        golden_sparse_utils()
        sys.exit(0)
    if "--only-prq" in sys.argv:
        golden_prq()
        golden_prmu()
        sys.exit(0)
    golden_sparse_utils()
    golden_prq()
    golden_prmu()
    golden_nbpc()
    golden_rates()
    golden_posterior()
    golden_estimation()
    golden_vae()
    golden_dataloader()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
