"""-m gpu: the public ``Posterior`` API of the product (``latents_map``, ``cell_noise_count_posterior_coo``,
``device_posterior``, ``compute_denoised_counts``) against the oracle restatement of the reference's host path
(reference posterior.py:424-594 COO loop, :669-782 noise_log_pdf / sample_mu_lambda_alpha, :861-917 latents map,
:282-330 denoised counts), on ONE device, with the latent draws injected on both sides.

Covered on purpose:
  * the lambda the reference's model() returns to the posterior is the one of its obs_empty block (epsilon = 1, y = 0;
    model.py:385-393 overwrites ``lam`` before the return at :445), y is forced to its MAP;
  * the window size is min(n_counts_max, max count of the reference's own 128-cell minibatch) (posterior.py:816), for a
    cell list whose size is not a multiple of the posterior batch size;
  * a trailing minibatch of fewer than 4 cells is dropped by the reference's loader (dataprep.py:155-161);
  * cells are visited in barcode order by the product, in loader order by the reference: same COO.
"""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from _helpers import _oracle_inputs, _setup
from oracle import estimation as oest
from oracle import nbpc as onb
from oracle import posterior as opost
from oracle import vae as ovae

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _trained(steps=6):
    ds, args, model, opt, svi, train_loader, _ = _setup()
    model.train()
    it = iter(train_loader)
    for _ in range(steps):
        try:
            svi.step(next(it))
        except StopIteration:
            it = iter(train_loader)
    torch.cuda.synchronize()
    model.eval()
    return ds, model


def _oracle_encode(ds, model, rows):
    """Oracle (float64) encoder outputs for analysed rows ``rows`` with the product's parameters, eval mode."""
    params, buffers, cfg = _oracle_inputs(model, dict(model.encoder["other"].offset))
    sd = {**params, **buffers}
    sub = lambda pre: {k[len(pre):]: v.detach() for k, v in sd.items() if k.startswith(pre)}
    x = torch.from_numpy(np.asarray(ds.get_count_matrix()[rows].todense(), dtype=np.float64))
    chi_ambient = model.param_store["chi_ambient"].detach().double().cpu()
    d_empty_loc = model.param_store["d_empty_loc"].detach().double().cpu()
    with torch.no_grad():
        zz = ovae.encode_z(x, sub("enc_z."))
        enc = ovae.encode_nonz(x, chi_ambient, zz["loc"], sub("enc_o."), d_empty_loc=d_empty_loc,
                               log_count_crossover=cfg["log_count_crossover"], prior_log_cell_counts=cfg["prior_log_cell_counts"],
                               empty_log_count_threshold=cfg["empty_log_count_threshold"], offset=cfg["offset"], training=False)
    return x, zz, enc, sub, cfg


@pytest.mark.parametrize("exact_gemm", [True, False])
def test_latents_map_matches_oracle(exact_gemm, request):
    """Posterior._get_latents_map (posterior.py:861-917): z, d, p, epsilon of every analysed droplet, 500-row batches,
    eval mode; float64 oracle with the same parameters.  With fp32 products 1e-4; as shipped (TF32) 1e-2 absolute on
    the logit-scale quantities."""
    from cellbender_b200.posterior import Posterior

    if exact_gemm:
        request.getfixturevalue("fp32_gemm")
    ds, model = _trained()
    post = Posterior(ds, model, posterior_batch_size=16)
    lat = post.latents_map
    n = ds.analyzed_barcode_inds.size
    assert lat["z"].shape == (n, model.z_dim) and lat["p"].shape == (n,) and lat["d"].shape == (n,)
    x, zz, enc, _, _ = _oracle_encode(ds, model, np.arange(n))
    ps = model.param_store
    d_ref = torch.exp(enc["d_loc"] + float(ps["d_cell_scale"]) ** 2 / 2).numpy()
    tol = 1e-4 if exact_gemm else 1e-2
    np.testing.assert_allclose(lat["z"], zz["loc"].numpy(), rtol=tol, atol=tol)
    np.testing.assert_allclose(lat["p"], enc["p_y"].sigmoid().numpy(), rtol=tol, atol=tol)
    np.testing.assert_allclose(lat["epsilon"], enc["epsilon"].numpy(), rtol=tol, atol=tol)
    np.testing.assert_allclose(lat["d"], d_ref, rtol=10 * tol)
    assert lat["phi_loc_scale"] == [pytest.approx(float(ps["phi_loc"])), pytest.approx(float(ps["phi_scale"]))]


@pytest.mark.parametrize("n_cells,pb", [(50, 16), (37, 16), (64, 16)])  # tails of 2 (dropped), 5 (kept), 0
def test_device_posterior_matches_reference_host_path(n_cells, pb, fp32_gemm):
    from cellbender_b200 import consts
    from cellbender_b200.posterior import Posterior

    ds, model = _trained()
    S, n_max, G = 5, 20, model.n_genes
    post = Posterior(ds, model, posterior_batch_size=pb, cells_per_chunk=24)  # chunks differ from the reference minibatches
    n = ds.analyzed_barcode_inds.size
    p = np.zeros(n)
    rng = np.random.default_rng(n_cells)
    cells = np.sort(rng.choice(np.arange(n), n_cells, replace=False))  # an arbitrary called set, loader (analysed) order
    p[cells] = 1.0
    post._latents = {"p": p, "z": None, "d": None, "epsilon": None, "phi_loc_scale": None}
    rows_sorted, n_window, barcodes = post.posterior_cell_order(n_max)
    tail = n_cells % pb
    n_eff = n_cells if (tail == 0 or tail >= consts.SMALLEST_ALLOWED_BATCH) else n_cells - tail
    assert rows_sorted.size == n_eff and np.all(np.diff(barcodes) > 0)
    assert set(rows_sorted.tolist()) == set(cells[:n_eff].tolist())

    # fixed latent draws per analysed row (phi per sample, shared by all chunks / minibatches)
    g = torch.Generator().manual_seed(3)
    Z = model.z_dim
    draws = {"phi": torch.rand(S, generator=g) * 0.3 + 0.05,
             "rho": {int(r): torch.rand(S, generator=g) * 0.1 for r in cells},
             "d_empty": {int(r): torch.exp(torch.randn(S, generator=g) * 0.05 + 4.0) for r in cells},
             "d_cell": {int(r): torch.exp(torch.randn(S, generator=g) * 0.1 + 7.0) for r in cells},
             "epsilon": {int(r): torch.rand(S, generator=g) * 0.4 + 0.8 for r in cells},
             "z": {int(r): torch.randn(S, Z, generator=g) for r in cells}}

    def noise_fn(pos, enc):
        rows = rows_sorted[pos]
        st = lambda k: torch.stack([draws[k][int(r)] for r in rows])
        return {"phi": draws["phi"], "rho": st("rho"), "d_empty": st("d_empty"), "d_cell": st("d_cell"),
                "epsilon": st("epsilon"), "z": st("z")}

    dev_post = post.device_posterior(n_samples=S, n_counts_max=n_max, noise_fn=noise_fn)
    m, c, lp = dev_post.m.cpu().numpy(), dev_post.c.cpu().numpy(), dev_post.log_prob.cpu().numpy()
    assert np.all(np.diff(m.astype(np.float64) * 64 + c) > 0), "COO must come out (m, c)-sorted without a sort"

    # ---- oracle: the reference's loop over ITS minibatches (loader order, batches of pb), float64 ----
    ref = {}
    off_ref = {}
    G_all = ds.matrix.shape[1]
    chi_ambient = model.param_store["chi_ambient"].detach().double().cpu()
    chi_bar = model.avg_gene_expression.double().cpu()
    for a in range(0, n_eff, pb):
        rows = cells[a:a + pb]
        x, zz, enc, sub, cfg = _oracle_encode(ds, model, rows)
        y = (enc["p_y"] > 0).double()
        samples = []
        for s in range(S):
            col = lambda k: torch.stack([draws[k][int(r)][s] for r in rows]).double()
            z = torch.stack([draws["z"][int(r)][s] for r in rows]).double()
            with torch.no_grad():
                chi = ovae.decode(z, sub("dec."), training=False)
            mu = onb.calculate_mu("full", epsilon=col("epsilon"), d_cell=col("d_cell"), chi=chi, y=y, rho=col("rho"))
            # the lambda model() returns: its obs_empty block (epsilon = 1, y = 0)  -- model.py:385-393, 445
            lam = onb.calculate_lambda("full", epsilon=torch.ones(len(rows), dtype=torch.float64), chi_ambient=chi_ambient,
                                       d_empty=col("d_empty"), y=torch.zeros(len(rows), dtype=torch.float64), d_cell=col("d_cell"),
                                       rho=col("rho"), chi_bar=chi_bar)
            samples.append((mu, lam, 1.0 / draws["phi"][s].double()))
        pdf, low = None, None
        for s, (mu, lam, al) in enumerate(samples, start=1):
            l, low = opost.log_prob_noise_count_tensor(x, mu + 1e-30, lam + 1e-30, al + 1e-30, n_counts_max=n_max,
                                                       mimic_float_casts=False)
            l = l - torch.logsumexp(l, -1, keepdim=True)
            pdf = l if s == 1 else torch.logaddexp(pdf + np.log(1 - 1 / s), l + np.log(1 / s))
        n_i, g_i, c_i, lp_i, off_i = opost.sparsify_posterior(pdf, low, x, -10.0)
        mm = opost.m_index(np.asarray(ds.analyzed_barcode_inds)[rows][n_i.numpy()], np.asarray(ds.analyzed_gene_inds)[g_i.numpy()], G_all)
        for k, cc, v, o in zip(mm, c_i.numpy(), lp_i.numpy(), off_i.numpy()):
            ref[(int(k), int(cc))] = float(v)
            if o != 0:
                off_ref[int(k)] = int(o)
    got = {(int(a_), int(b_)): float(v) for a_, b_, v in zip(m, c, lp)}
    edge = lambda v: abs(v + 10.0) < 2e-3
    assert len(got) > 100
    assert all(k in got or edge(v) for k, v in ref.items()), "entries of the reference COO are missing"
    assert all(k in ref or edge(v) for k, v in got.items()), "entries the reference does not have"
    common = [k for k in ref if k in got]
    err = np.array([abs(ref[k] - got[k]) / max(abs(ref[k]), 1.0) for k in common])
    assert err.max() < 2e-4, err.max()
    off_got = dev_post.offsets_dict()
    for k, v in off_got.items():
        assert off_ref.get(k) == v

    # ---- the scipy-returning public method and the estimator on top of it ----
    post._device_posterior = dev_post
    shape = [int(np.prod(ds.matrix.shape, dtype=np.uint64)), n_max]
    coo = dev_post.to_scipy(shape)
    assert coo.shape == tuple(shape) and coo.nnz == len(got)
    from cellbender_b200.estimation import MAP

    denoised = post.compute_denoised_counts(MAP, device=DEV)
    assert sp.issparse(denoised) and denoised.shape == ds.matrix.shape and denoised.format == "csc"
    noise_ref = oest.estimate_map(coo, off_got, ds.matrix.shape)
    cell_inds = np.asarray(ds.analyzed_barcode_inds)[p > 0.5]
    keep = np.zeros(ds.matrix.shape[0], dtype=bool)
    keep[cell_inds] = True
    want = (sp.csr_matrix(sp.diags(keep.astype(ds.matrix.dtype)) @ ds.matrix) - noise_ref).tocsc()
    want.eliminate_zeros()
    assert (denoised != want).nnz == 0
    assert np.all(denoised.data >= 0)


def test_public_posterior_coo_caches_and_sums_to_one():
    """cell_noise_count_posterior_coo(): scipy COO of shape [N_all * G_all, n_counts_max], cached per kwargs
    (posterior.py:388-420); every (cell, gene) row is a normalised pmf up to the -10 threshold."""
    from cellbender_b200.posterior import Posterior

    ds, model = _trained()
    post = Posterior(ds, model, posterior_batch_size=16)
    n = ds.analyzed_barcode_inds.size
    p = np.zeros(n)
    p[:40] = 1.0
    post._latents = {"p": p, "z": None, "d": None, "epsilon": None, "phi_loc_scale": None}
    coo = post.cell_noise_count_posterior_coo(n_samples=3)
    assert coo is post.cell_noise_count_posterior_coo(n_samples=3)
    assert coo.shape == (ds.matrix.shape[0] * ds.matrix.shape[1], 20)
    tot = np.zeros(0)
    rows, inv = np.unique(coo.row, return_inverse=True)
    tot = np.bincount(inv, weights=np.exp(coo.data.astype(np.float64)))
    assert tot.max() <= 1.0 + 1e-5 and tot.min() > 0.99  # at most exp(-10) * 20 is thresholded away
    raw = ds.matrix.tocsr()
    nn, gg = np.divmod(rows, ds.matrix.shape[1])
    assert np.all(np.asarray(raw[nn, gg]).ravel() > 0), "posterior rows exist only where the count is non-zero"
    with pytest.raises(RuntimeError, match="Zero cells found"):
        post2 = Posterior(ds, model)
        post2._latents = {"p": np.zeros(n)}
        post2.cell_noise_count_posterior_coo()


@pytest.mark.parametrize("n_called", [100, 70])
def test_graph_replayed_chunks_equal_eager_chunks(monkeypatch, n_called):
    """Full chunks of a posterior pass replay ONE captured CUDA graph (static operand buffers, device-side output
    offset, grid sized for the largest chunk); the ragged last chunk runs eagerly.  With the Monte-Carlo draws replaced
    by their means on both paths the two must produce the identical COO."""
    from cellbender_b200.posterior import Posterior

    ds, model = _trained()
    monkeypatch.setattr(torch, "randn", lambda shape, device=None, **k: torch.zeros(shape, device=device))
    monkeypatch.setattr(torch, "_standard_gamma", lambda conc: conc.clone())
    n = ds.analyzed_barcode_inds.size
    p = np.zeros(n)
    p[:n_called] = 1.0  # 100: 3 full chunks + 4 cells (three-deep pipeline of graphs); 70: 2 full chunks + 6 (one graph per chunk)
    outs = []
    for use_graph in (True, False):
        post = Posterior(ds, model, posterior_batch_size=16, cells_per_chunk=32)  # 3 full chunks + a tail of 4 cells
        post.use_cuda_graph = use_graph
        post._latents = {"p": p, "z": None, "d": None, "epsilon": None, "phi_loc_scale": None}
        a = post.device_posterior(n_samples=4)
        b = post.device_posterior(n_samples=4)  # second pass: the cached graph and workspace are reused
        assert (getattr(post, "_chunk_graph" if n_called == 100 else "_chunk_graph_serial", None) is not None) == use_graph
        for x, y in zip((a.m, a.c, a.log_prob, a.off_m, a.off_v), (b.m, b.c, b.log_prob, b.off_m, b.off_v)):
            assert torch.equal(x, y)
        outs.append(a)
    g, e = outs
    assert g.m.numel() > 700
    assert torch.equal(g.m, e.m) and torch.equal(g.c, e.c) and torch.equal(g.off_m, e.off_m) and torch.equal(g.off_v, e.off_v)
    torch.testing.assert_close(g.log_prob, e.log_prob, rtol=0, atol=0)
