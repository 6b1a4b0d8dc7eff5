"""-m gpu parity tests: every CUDA kernel, called through the C ABI, against the oracle / golden vectors."""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import estimation as oest
from oracle import nbpc as onb
from oracle import posterior as opost

pytestmark = pytest.mark.gpu

DEV = "cuda"


@pytest.fixture(scope="module")
def ops():
    from cellbender_b200 import build, ops as _ops

    build.build()
    return _ops


def _rand_csr(rng, R, G, density, scale=3.0):
    dense = rng.poisson(scale, size=(R, G)) * (rng.random((R, G)) < density)
    return sp.csr_matrix(dense.astype(np.float32)), dense.astype(np.float32)


# ---------------------------------------------------------------------------------------------- gather
@pytest.mark.parametrize("G", [30, 33, 1000])
def test_gather_matches_reference_collate(ops, G):
    rng = np.random.default_rng(0)
    csr, dense = _rand_csr(rng, 40, G, 0.2)
    dense[3] = 0  # an empty droplet row
    csr = sp.csr_matrix(dense)
    d = ops.DeviceCSR.from_scipy(csr, DEV)
    rows = np.array([5, 3, 3, 39, 0, 17], dtype=np.int64)
    amb = rng.random(G).astype(np.float32)
    amb /= np.linalg.norm(amb)
    out = ops.gather_rows(d, torch.as_tensor(rows, device=DEV), torch.as_tensor(amb, device=DEV))
    x = dense[rows]
    np.testing.assert_array_equal(out["raw"][:, :G].cpu().numpy(), x)  # bit-exact: integer counts
    t = torch.tensor(x)
    xn = t / (t.sum(-1, keepdim=True) + 1e-5)
    np.testing.assert_array_equal(out["norm"][:, :G].cpu().numpy(), xn.numpy())  # same fp32 division
    np.testing.assert_allclose(out["lognorm"][:, :G].cpu().numpy(), xn.log1p().numpy(), rtol=2e-7, atol=0)
    f = out["feats"].cpu().numpy()
    np.testing.assert_array_equal(f[:, 0], x.sum(-1))
    np.testing.assert_array_equal(f[:, 1], (x > 0).sum(-1))
    np.testing.assert_allclose(f[:, 2], x @ amb, rtol=1e-5)
    np.testing.assert_array_equal(f[:, 3], x.max(-1))


def test_gather_full_size_roundtrip(ops):
    """BASELINE size (255 x 33538): scatter -> dense -> nonzero round trip is exact, padding stays zero."""
    rng = np.random.default_rng(1)
    G, R = 33538, 600
    nnz_per_row = rng.integers(20, 3000, size=R)
    indptr = np.concatenate([[0], np.cumsum(nnz_per_row)]).astype(np.int64)
    indices = np.concatenate([np.sort(rng.choice(G, size=k, replace=False)) for k in nnz_per_row]).astype(np.int32)
    data = rng.integers(1, 50, size=indices.size).astype(np.float32)
    csr = sp.csr_matrix((data, indices, indptr), shape=(R, G))
    d = ops.DeviceCSR.from_scipy(csr, DEV)
    rows = rng.integers(0, R, size=255).astype(np.int64)
    out = ops.gather_rows(d, torch.as_tensor(rows, device=DEV), None, want=("raw",))
    raw = out["raw"]
    assert raw.shape[1] % 4 == 0 and float(raw[:, G:].abs().sum()) == 0.0
    back = sp.csr_matrix(raw[:, :G].cpu().numpy())
    ref = csr[rows]
    assert (back != ref).nnz == 0
    np.testing.assert_array_equal(out["feats"][:, 0].cpu().numpy(), np.asarray(ref.sum(axis=1)).ravel())


# ------------------------------------------------------------------------------------------- likelihood
def test_nbpc_elementwise_vs_multiprecision(ops, golden_dir):
    g = np.load(os.path.join(golden_dir, "nbpc.npz"))
    v, mu, al, lam = (torch.tensor(g[k], dtype=torch.float32, device=DEV) for k in ("value", "mu", "alpha", "lam"))
    mu.requires_grad_(True), al.requires_grad_(True), lam.requires_grad_(True)
    lp = ops.nbpc_approx_log_prob(v, mu, al, lam)
    ref = g["approx_mp"]
    err = np.abs(lp.detach().cpu().numpy() - ref) / np.maximum(np.abs(ref), 1.0)
    assert err.max() < 1e-5, err.max()  # every count: above 64 the element takes the float64 pipe (nbpc_eval_large_f64)
    gm, ga, gl = torch.autograd.grad(lp.sum(), (mu, al, lam))
    for name, mine in (("dmu", gm), ("dalpha", ga), ("dlam", gl)):
        r = g[f"approx_{name}_mp"]
        e = np.abs(mine.cpu().numpy() - r) / np.maximum(np.abs(r), 1e-3)
        assert e.max() < 1e-3 and np.mean(e > 1e-4) < 0.01, (name, e.max())


def test_nbpc_broadcast_like_pyro_enumeration(ops):
    """value [B,G], mu/lam [2,B,G], scalar alpha: the shapes at model.py:338-346; gradient reductions follow."""
    rng = np.random.default_rng(2)
    B, G = 5, 37
    value = torch.tensor(rng.poisson(2.0, (B, G)).astype(np.float32), device=DEV)
    mu = torch.tensor(rng.gamma(2.0, 1.0, (2, B, G)).astype(np.float32) + 1e-10, device=DEV, requires_grad=True)
    lam = torch.tensor(rng.gamma(1.0, 0.3, (2, B, G)).astype(np.float32) + 1e-10, device=DEV, requires_grad=True)
    alpha = torch.tensor(4.3, device=DEV, requires_grad=True)
    lp = ops.nbpc_approx_log_prob(value, mu, alpha, lam)
    assert lp.shape == (2, B, G)
    wts = torch.tensor(rng.random((2, B, G)).astype(np.float32), device=DEV)
    gm, ga, gl = torch.autograd.grad((lp * wts).sum(), (mu, alpha, lam))
    L, dmu, dlam, dal = onb.nbpc_approx_mp(value.cpu().numpy()[None], mu.detach().cpu().numpy(),
                                           float(alpha), lam.detach().cpu().numpy())
    np.testing.assert_allclose(lp.detach().cpu().numpy(), L, rtol=1e-5, atol=1e-5)
    w = wts.cpu().numpy()
    np.testing.assert_allclose(gm.cpu().numpy(), dmu * w, rtol=2e-4, atol=1e-5)
    np.testing.assert_allclose(gl.cpu().numpy(), dlam * w, rtol=2e-4, atol=1e-5)
    assert ga.shape == () and abs(float(ga) - float((dal * w).sum())) < 2e-4 * max(1.0, abs(float((dal * w).sum())))


def test_nbpc_raises_nan_exception(ops):
    from cellbender_b200.exceptions import NanException

    v = torch.zeros(4, device=DEV)
    mu = torch.tensor([1.0, float("nan"), 1.0, 1.0], device=DEV)
    with pytest.raises(NanException):
        ops.nbpc_approx_log_prob(v, mu, torch.tensor(2.0, device=DEV), torch.ones(4, device=DEV))


def _obs_case(rng, B, G, density=0.15, big_counts=False):
    dense = rng.poisson(2.5, size=(B, G)) * (rng.random((B, G)) < density)
    if big_counts:
        dense[0, :3] = [40, 180, 700]
    dense[-1] = 0  # droplet with no counts at all
    dense = dense.astype(np.float32)
    logits = (rng.normal(size=(B, G)) * 2.5).astype(np.float32)
    chi_a = rng.dirichlet(np.full(G, 0.3)).astype(np.float32) + 1e-12
    chi_b = rng.dirichlet(np.full(G, 0.5)).astype(np.float32) + 1e-12
    eps = rng.gamma(50, 1 / 50, B)
    d_cell = np.exp(rng.normal(np.log(3000), 0.3, B))
    d_empty = np.exp(rng.normal(np.log(100), 0.2, B))
    rho = rng.beta(3, 80, B)
    coef = np.stack([(1 - rho) * eps * d_cell, eps * (1 - rho) * d_empty, eps * rho * (d_cell + d_empty),
                     eps * (1 - rho) * d_empty, eps * rho * d_empty, (1 - rho) * d_empty, rho * d_empty], 1)
    w = -np.stack([rng.random(B), rng.random(B), 0.01 * rng.random(B)], 1)
    return dense, logits, chi_a, chi_b, coef.astype(np.float32), np.float32(5.3), w.astype(np.float32)


@pytest.mark.parametrize("B,G,big", [(6, 96, True), (5, 131, False), (3, 1, False)])
def test_elbo_obs_fused_vs_multiprecision_oracle(ops, B, G, big):
    rng = np.random.default_rng(B * 1000 + G)
    dense, logits, chi_a, chi_b, coef, alpha, w = _obs_case(rng, B, G, big_counts=big and G > 3)
    csr_all = sp.csr_matrix(np.vstack([dense, dense[::-1]]))  # row_ids pick a permutation out of a larger matrix
    rows = np.arange(B, dtype=np.int64)[::-1].copy() + B      # -> rows equal to dense[0..B-1]
    d = ops.DeviceCSR.from_scipy(csr_all, DEV)
    lg = ops.alloc_rows(B, G, DEV)
    lg[:, :G] = torch.tensor(logits)
    ld = lg.shape[1]
    ca, cb = torch.zeros(ld, device=DEV), torch.zeros(ld, device=DEV)
    ca[:G], cb[:G] = torch.tensor(chi_a), torch.tensor(chi_b)
    out = ops.elbo_obs(d, torch.as_tensor(rows, device=DEV), lg, ca, cb, torch.tensor(coef, device=DEV),
                       torch.tensor([alpha], device=DEV), torch.tensor(w, device=DEV))
    ref = onb.elbo_obs_oracle(dense, logits, chi_a, chi_b, coef, alpha, w)
    assert int(out["nan_flag"].item()) == 0
    np.testing.assert_allclose(out["lse"].cpu().numpy(), ref["lse"], rtol=1e-6, atol=1e-6)
    sums = out["sums"].cpu().numpy().astype(np.float64)
    # log-prob row sums: 1e-5 relative (north_star); gradient sums: 1e-4 of their scale
    np.testing.assert_allclose(sums[:, :3], ref["sums"][:, :3], rtol=1e-5, atol=1e-5)
    scale = np.abs(ref["sums"][:, 3:]).max(0, keepdims=True) + 1e-3
    assert (np.abs(sums[:, 3:] - ref["sums"][:, 3:]) / scale).max() < 2e-4
    dl = out["dlogits"][:, :G].cpu().numpy()
    assert np.abs(dl - ref["dlogits"]).max() < 2e-4 * (np.abs(ref["dlogits"]).max() + 1e-6)
    dca = out["dchi_ambient"].cpu().numpy()
    assert np.abs(dca - ref["dchi_ambient"]).max() < 2e-4 * (np.abs(ref["dchi_ambient"]).max() + 1e-6)


def test_elbo_obs_full_size_properties(ops):
    """PBMC8k-shape minibatch (255 x 33538): properties that hold at any size.
    (1) softmax backward: each dlogits row sums to 0; (2) the fused row sums equal the sums of the
    elementwise kernel over the same rates; (3) the dense/sparse split covers every gene exactly once."""
    rng = np.random.default_rng(7)
    B, G = 255, 33538
    nnz_per_row = np.concatenate([rng.integers(800, 3000, size=204), rng.integers(5, 150, size=51)])
    indptr = np.concatenate([[0], np.cumsum(nnz_per_row)]).astype(np.int64)
    indices = np.concatenate([np.sort(rng.choice(G, size=k, replace=False)) for k in nnz_per_row]).astype(np.int32)
    data = (rng.geometric(0.5, size=indices.size)).astype(np.float32)
    d = ops.DeviceCSR.from_scipy(sp.csr_matrix((data, indices, indptr), shape=(B, G)), DEV)
    rows = torch.arange(B, device=DEV)
    lg = ops.alloc_rows(B, G, DEV)
    lg[:, :G] = torch.randn(B, G, device=DEV) * 2.0
    ld = lg.shape[1]
    chi_a = torch.zeros(ld, device=DEV)
    chi_b = torch.zeros(ld, device=DEV)
    chi_a[:G] = torch.softmax(torch.randn(G, device=DEV) * 2, 0)
    chi_b[:G] = torch.softmax(torch.randn(G, device=DEV) * 2, 0)
    _, _, _, _, coef, alpha, w = _obs_case(rng, B, 8)
    coef_t, w_t, al_t = torch.tensor(coef, device=DEV), torch.tensor(w, device=DEV), torch.tensor([alpha], device=DEV)
    out = ops.elbo_obs(d, rows, lg, chi_a, chi_b, coef_t, al_t, w_t)
    torch.cuda.synchronize()
    assert int(out["nan_flag"].item()) == 0
    dl = out["dlogits"][:, :G]
    assert float(dl.sum(-1).abs().max()) < 1e-3 * float(dl.abs().sum(-1).max())
    # (2) elementwise kernel on the materialised rates
    x = ops.gather_rows(d, rows, None, want=("raw",))["raw"][:, :G]
    chi = torch.softmax(lg[:, :G], -1)
    mu1 = coef_t[:, 0:1] * chi + 1e-10
    lam1 = coef_t[:, 1:2] * chi_a[:G] + coef_t[:, 2:3] * chi_b[:G] + 1e-10
    lam0 = coef_t[:, 3:4] * chi_a[:G] + coef_t[:, 4:5] * chi_b[:G] + 1e-10
    L1 = ops.nbpc_approx_log_prob(x, mu1, al_t[0], lam1).double().sum(-1)
    L0 = ops.nbpc_approx_log_prob(x, torch.full_like(mu1, 1e-10), al_t[0], lam0).double().sum(-1)
    s = out["sums"].double()
    np.testing.assert_allclose(s[:, 0].cpu().numpy(), L1.cpu().numpy(), rtol=1e-5)
    np.testing.assert_allclose(s[:, 1].cpu().numpy(), L0.cpu().numpy(), rtol=1e-5)
    # (3) LE = sum_g [x log lamE - lamE - lgamma(x+1)] evaluated independently in fp64 torch
    lamE = (coef_t[:, 5:6] * chi_a[:G] + coef_t[:, 6:7] * chi_b[:G] + 1e-10).double()
    LE = (torch.xlogy(x.double(), lamE) - lamE - torch.lgamma(x.double() + 1)).sum(-1)
    np.testing.assert_allclose(s[:, 2].cpu().numpy(), LE.cpu().numpy(), rtol=1e-5)


# ------------------------------------------------------------------------------------------ small utils
def test_exclusive_scan_and_colsum_and_lse(ops):
    x = torch.randint(0, 5, (1_000_003,), dtype=torch.int32, device=DEV)
    pos, total = ops.exclusive_scan_i32(x)
    ref = torch.cumsum(x.long(), 0) - x.long()
    assert torch.equal(pos, ref) and int(total.item()) == int(x.long().sum().item())
    m = ops.alloc_rows(37, 1001, DEV)
    m[:, :1001] = torch.randn(37, 1001, device=DEV)
    np.testing.assert_allclose(ops.colsum_rows(m, 1001).cpu().numpy(), m[:, :1001].double().sum(0).cpu().numpy(),
                               rtol=1e-5, atol=1e-5)
    add = torch.randn(37, device=DEV)
    np.testing.assert_allclose(ops.col_logsumexp(m, 37, 1001, row_add=add).cpu().numpy(),
                               torch.logsumexp(m[:, :1001].double() + add.double()[:, None], 0).cpu().numpy(), rtol=1e-6, atol=1e-6)
    big = torch.randn(5000, 300, device=DEV) * 3
    np.testing.assert_allclose(ops.col_logsumexp(big, 5000, 300).cpu().numpy(),
                               torch.logsumexp(big.double(), 0).cpu().numpy(), rtol=1e-6, atol=1e-6)


def test_clipped_adam_matches_per_tensor_reference(ops):
    """Fused flat ClippedAdam + OneCycle tables vs a plain per-tensor restatement of pyro's ClippedAdam driven by
    torch's own OneCycleLR (run.py:593-629)."""
    torch.manual_seed(0)
    n, steps, lr = 10_007, 7, 1e-4
    p0 = torch.randn(n, dtype=torch.float64)
    grads = [torch.randn(n, dtype=torch.float64) * (30.0 if k == 2 else 1.0) for k in range(steps)]  # exercise the clamp
    # reference: torch OneCycleLR on a dummy Adam to obtain lr / beta1 per step, ClippedAdam update in fp64
    dummy = torch.optim.Adam([torch.zeros(1, requires_grad=True)], lr=lr, betas=(0.9, 0.999))
    sched = torch.optim.lr_scheduler.OneCycleLR(dummy, max_lr=lr * 10, total_steps=steps)
    lrs, b1s = [], []
    p, m, v = p0.clone(), torch.zeros(n, dtype=torch.float64), torch.zeros(n, dtype=torch.float64)
    for k in range(steps):
        lr_k, b1 = dummy.param_groups[0]["lr"], dummy.param_groups[0]["betas"][0]
        lrs.append(lr_k), b1s.append(b1)
        g = grads[k].clamp(-10.0, 10.0)
        m = b1 * m + (1 - b1) * g
        v = 0.999 * v + 0.001 * g * g
        t = k + 1
        p = p - lr_k * np.sqrt(1 - 0.999 ** t) / (1 - b1 ** t) * m / (v.sqrt() + 1e-8)
        dummy.step()
        if k < steps - 1:
            sched.step()
    pd = p0.float().to(DEV)
    md, vd = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    step = torch.zeros(1, dtype=torch.int64, device=DEV)
    lr_t, b1_t = torch.tensor(lrs, dtype=torch.float32, device=DEV), torch.tensor(b1s, dtype=torch.float32, device=DEV)
    for k in range(steps):
        ops.clipped_adam_step(pd, grads[k].float().to(DEV), md, vd, lr_t, b1_t, step)
    assert int(step.item()) == steps
    # the same sweep with the host-built step-size table (what optim.FlatClippedAdam passes): identical to 1e-7
    t_ = np.arange(1, steps + 1, dtype=np.float64)
    st_t = torch.tensor(lr_t.double().cpu().numpy() * np.sqrt(1 - np.float64(np.float32(0.999)) ** t_)
                        / (1 - b1_t.double().cpu().numpy() ** t_), dtype=torch.float32, device=DEV)
    pd2, md2, vd2 = p0.float().to(DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    step2 = torch.zeros(1, dtype=torch.int64, device=DEV)
    for k in range(steps):
        ops.clipped_adam_step(pd2, grads[k].float().to(DEV), md2, vd2, lr_t, b1_t, step2, step_table=st_t)
    torch.testing.assert_close(pd2, pd, rtol=1e-6, atol=1e-7)
    assert torch.equal(md2, md) and torch.equal(vd2, vd)
    np.testing.assert_allclose(pd.cpu().numpy(), p.numpy(), rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(md.cpu().numpy(), m.numpy(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(vd.cpu().numpy(), v.numpy(), rtol=1e-4, atol=1e-7)


# ------------------------------------------------------------------------------------------- estimators
def _case(g, tag):
    nc, ng, ncol = (int(v) for v in g[f"{tag}_shape"])
    coo = sp.coo_matrix((g[f"{tag}_data"], (g[f"{tag}_row"], g[f"{tag}_col"])), shape=(nc * ng, ncol))
    off = dict(zip(g[f"{tag}_off_m"].tolist(), g[f"{tag}_off_v"].tolist()))
    return coo, off, (nc, ng)


CASES = ["fixture"] + [f"rand{i}" for i in range(6)]


@pytest.mark.parametrize("tag", CASES)
def test_estimators_bit_exact_vs_reference_golden(ops, golden_dir, tag):
    from cellbender_b200 import estimation as est

    g = np.load(os.path.join(golden_dir, "estimation.npz"))
    coo, off, shape = _case(g, tag)
    conv = est.IndexConverter(*shape)
    np.testing.assert_array_equal(np.array(est.MAP(conv).estimate_noise(coo, off).todense()), g[f"{tag}_map"])
    np.testing.assert_array_equal(np.array(est.ThresholdCDF(conv).estimate_noise(coo, off, q=0.5).todense()),
                                  g[f"{tag}_cdf50"])
    np.testing.assert_array_equal(np.array(est.ThresholdCDF(conv).estimate_noise(coo, off, q=0.9).todense()),
                                  g[f"{tag}_cdf90"])
    mean = est.Mean(conv).estimate_noise(coo, off)
    assert mean.dtype == np.float32
    np.testing.assert_allclose(np.array(mean.todense()), g[f"{tag}_mean"], rtol=1e-6, atol=1e-7)
    for tname in ("t0", "tmap", "tplus", "tminus", "tmix", "tbig"):
        out = est.MultipleChoiceKnapsack(conv).estimate_noise(coo, off, noise_targets_per_gene=g[f"{tag}_mckp_{tname}_target"])
        assert out.dtype == np.int32 and out.shape == shape
        np.testing.assert_array_equal(np.array(out.todense()), g[f"{tag}_mckp_{tname}"], err_msg=f"{tag}/{tname}")
    # SingleSample: the sample lies in the support (the reference's own criterion, tests/test_estimation.py:159-188)
    smp = np.array(est.SingleSample(conv).estimate_noise(coo, off, seed=3).todense()).ravel()
    dense = np.full((shape[0] * shape[1], coo.shape[1]), -np.inf)
    dense[coo.row, coo.col] = coo.data
    for m_ in np.unique(coo.row):
        allowed = np.flatnonzero(dense[m_] > -np.inf) + off.get(int(m_), 0)
        assert smp[m_] in allowed
    # target function (posterior.py:1589-1650)
    raw = sp.csr_matrix(g[f"{tag}_raw"].astype(np.int32))
    fn = est.compute_mean_target_removal_as_function(coo, off, conv, raw, n_cells=shape[0], device=DEV, per_gene=True)
    np.testing.assert_allclose(fn(0.01).cpu().numpy(), g[f"{tag}_target_fpr01"], rtol=1e-6, atol=1e-7)


def test_mckp_reference_truth_tables(ops, golden_dir):
    """The reference's own known answers (tests/test_estimation.py:289-305), all three COO variants."""
    from cellbender_b200 import estimation as est

    g = np.load(os.path.join(golden_dir, "estimation.npz"))
    coo, off, _ = _case(g, "fixture")
    rng = np.random.default_rng(0)
    perm = rng.permutation(coo.nnz)
    keep = coo.data >= -6
    variants = {"exact": coo,
                "filtered": sp.coo_matrix((coo.data[keep], (coo.row[keep], coo.col[keep])), shape=coo.shape),
                "unsorted": sp.coo_matrix((coo.data[perm], (coo.row[perm], coo.col[perm])), shape=coo.shape)}
    table = [(1, np.zeros(8), [0, 1, 2, 0, 0, 0, 0, 1], None), (1, np.ones(8), [0, 1, 2, 1, 1, 1, 1, 1], None),
             (1, np.ones(8) * 2, [0, 1, 2, 2, 2, 2, 2, 2], None), (4, np.zeros(2), [2, 2], None),
             (4, np.ones(2) * 4, [4, 4], [[0, 1], [2, 2], [2, 0], [0, 1]]),
             (4, np.ones(2) * 9, [9, 9], [[0, 1], [2, 3], [4, 2], [3, 3]])]
    for name, c in variants.items():
        for n_cells, target, truth, truth_mat in table:
            conv = est.IndexConverter(n_cells, 8 // n_cells)
            out = np.array(est.MultipleChoiceKnapsack(conv).estimate_noise(c, off, noise_targets_per_gene=target,
                                                                           n_chunks=2).todense())
            if truth_mat is not None:
                np.testing.assert_array_equal(out, np.array(truth_mat), err_msg=name)
            np.testing.assert_array_equal(out.sum(axis=0), np.array(truth), err_msg=name)


def test_estimators_large_random_vs_oracle(ops):
    """A posterior of realistic structure (2e5 entries, m beyond int32) against the oracle restatement."""
    from cellbender_b200 import estimation as est

    rng = np.random.default_rng(5)
    n_cells, n_genes = 70_000, 33_538  # n_cells * n_genes > 2^31
    cells = np.sort(rng.choice(n_cells, 300, replace=False))
    rows, cols, vals, off = [], [], [], {}
    for n in cells:
        for g_ in np.sort(rng.choice(n_genes, 60, replace=False)):
            m = int(n) * n_genes + int(g_)
            k = int(rng.integers(1, 12))
            lp = rng.normal(size=k) * 2
            lp -= np.log(np.exp(lp).sum())
            keep = lp >= -6
            keep[np.argmax(lp)] = True
            for c in np.flatnonzero(keep):
                rows.append(m), cols.append(c), vals.append(lp[c])
            if rng.random() < 0.2:
                off[m] = int(rng.integers(1, 5))
    coo = sp.coo_matrix((np.array(vals, np.float32), (np.array(rows, np.int64), np.array(cols, np.int32))),
                        shape=(n_cells * n_genes, 12))
    conv = est.IndexConverter(n_cells, n_genes)
    shape = (n_cells, n_genes)
    assert (est.MAP(conv).estimate_noise(coo, off) != oest.estimate_map(coo, off, shape)).nnz == 0
    assert (est.ThresholdCDF(conv).estimate_noise(coo, off, q=0.7) != oest.estimate_cdf(coo, off, shape, q=0.7)).nnz == 0
    map_per_gene = np.asarray(oest.estimate_map(coo, off, shape).sum(axis=0)).ravel().astype(float)
    target = map_per_gene + rng.integers(-3, 4, size=n_genes) + 0.25
    got = est.MultipleChoiceKnapsack(conv).estimate_noise(coo, off, noise_targets_per_gene=target)
    assert (got != oest.estimate_mckp(coo, off, shape, target)).nnz == 0


def test_estimator_errors_match_reference(ops):
    from cellbender_b200 import estimation as est

    conv = est.IndexConverter(2, 4)
    coo = sp.coo_matrix((np.array([0.0], np.float32), (np.array([1]), np.array([0]))), shape=(8, 3))
    with pytest.raises(AssertionError):
        est.MultipleChoiceKnapsack(conv).estimate_noise(coo, {}, noise_targets_per_gene=np.zeros(3))
    with pytest.raises(AssertionError):
        est.MultipleChoiceKnapsack(conv).estimate_noise(coo, {}, noise_targets_per_gene=None)
    with pytest.raises(ValueError):
        conv.get_ng_indices(np.array([99]))
    bad = sp.coo_matrix((np.array([0.0], np.float32), (np.array([9]), np.array([0]))), shape=(10, 3))
    with pytest.raises(ValueError):
        est.MAP(conv).estimate_noise(bad, {})


# -------------------------------------------------------------------------------------------- posterior
@pytest.mark.parametrize("scale,n_max", [(30.0, 20), (2.0, 20), (300.0, 20), (30.0, 32)])
def test_posterior_coo_vs_dense_oracle(ops, scale, n_max):
    """cb_posterior_* against the dense float64 restatement of posterior.py:669-859 + :528-576."""
    rng = np.random.default_rng(int(scale) + n_max)
    Bc, G, S = 7, 61, 5
    G_all, N_all = 80, 50
    gene_all = np.sort(rng.choice(G_all, G, replace=False)).astype(np.int64)
    barcode_all = np.sort(rng.choice(N_all, Bc, replace=False)).astype(np.int64)
    logits = (rng.normal(size=(S * Bc, G)) * 2).astype(np.float32)
    chi = torch.softmax(torch.tensor(logits, dtype=torch.float64), -1)
    chi_a = rng.dirichlet(np.full(G, 0.5)).astype(np.float32)
    chi_b = rng.dirichlet(np.full(G, 0.5)).astype(np.float32)
    a_mu = (np.exp(rng.normal(np.log(scale * G), 0.3, S * Bc))).astype(np.float32)
    a_mu[1] = 0.0  # a y = 0 droplet: mu = 0 + 1e-30
    c_a = (np.exp(rng.normal(np.log(scale * G * 0.1), 0.3, S * Bc))).astype(np.float32)
    c_b = (c_a * 0.05).astype(np.float32)
    alpha = rng.gamma(5.0, 1.0, S).astype(np.float32)
    mu_all = torch.tensor(a_mu, dtype=torch.float64)[:, None] * chi
    lam_all = torch.tensor(c_a, dtype=torch.float64)[:, None] * torch.tensor(chi_a, dtype=torch.float64) \
        + torch.tensor(c_b, dtype=torch.float64)[:, None] * torch.tensor(chi_b, dtype=torch.float64)
    data = torch.poisson(mu_all[:Bc] + lam_all[:Bc], generator=torch.Generator().manual_seed(1))
    data[torch.rand(Bc, G, generator=torch.Generator().manual_seed(2)) < 0.5] = 0.0
    samples = [(mu_all[s * Bc:(s + 1) * Bc], lam_all[s * Bc:(s + 1) * Bc], torch.tensor(float(alpha[s]), dtype=torch.float64))
               for s in range(S)]
    # oracle (true float64; the window size n is the batch-wide min(n_max, max count))
    pdf, low = None, None
    for s, (mu, lam, al) in enumerate(samples, start=1):
        lp, low = opost.log_prob_noise_count_tensor(data, mu + 1e-30, lam + 1e-30, al + 1e-30, n_counts_max=n_max,
                                                    mimic_float_casts=False)
        lp = lp - torch.logsumexp(lp, -1, keepdim=True)
        pdf = lp if s == 1 else torch.logaddexp(pdf + np.log(1 - 1 / s), lp + np.log(1 / s))
    n_i, g_i, c_i, lp_i, off_i = opost.sparsify_posterior(pdf, low, data, -10.0)
    m_ref = opost.m_index(barcode_all[n_i.numpy()], gene_all[g_i.numpy()], G_all)
    # device: transposed, cell-major logits [G, Bc * S] (column b * S + s), decoder bias passed separately
    d = ops.DeviceCSR.from_scipy(sp.csr_matrix(data.numpy().astype(np.float32)), DEV)
    bias = (rng.normal(size=G) * 0.5).astype(np.float32)
    cm = lambda v: np.ascontiguousarray(np.asarray(v).reshape(S, Bc).T).reshape(-1)  # sample-major [S * Bc] -> cell-major [Bc * S]
    lt = ops.alloc_rows(G, Bc * S, DEV)
    lt[:, :Bc * S] = torch.tensor(np.ascontiguousarray((logits - bias[None, :]).reshape(S, Bc, G).transpose(2, 1, 0)).reshape(G, Bc * S))
    t = lambda a, dt=torch.float32: torch.as_tensor(a, dtype=dt, device=DEV)
    lse = ops.col_logsumexp(lt, G, Bc * S, row_add=t(bias))
    lse_ref = torch.logsumexp(torch.tensor(logits, dtype=torch.float64), -1).numpy()
    np.testing.assert_allclose(lse.cpu().numpy(), cm(lse_ref), rtol=2e-6, atol=2e-6)
    n_win = int(min(n_max, data.max().item()))
    ld = ops.round_up(G, 4)
    ca, cb = torch.zeros(ld, device=DEV), torch.zeros(ld, device=DEV)
    ca[:G], cb[:G] = t(chi_a), t(chi_b)
    m, c, lp, off_m, off_v = ops.posterior_coo(
        d, torch.arange(Bc, device=DEV), lt, t(bias), lse, ca, cb, t(cm(a_mu)), t(cm(c_a)), t(cm(c_b)), t(alpha),
        torch.full((Bc,), n_win, dtype=torch.int32, device=DEV), t(barcode_all, torch.int64), t(gene_all, torch.int64),
        G_all, n_counts_max=n_max)
    m, c, lp = m.cpu().numpy(), c.cpu().numpy(), lp.cpu().numpy()
    # entries whose float64 log prob sits within 1e-4 of the -10 threshold may legitimately fall on either side
    ref = {(int(a), int(b)): float(v) for a, b, v in zip(m_ref, c_i.numpy(), lp_i.numpy())}
    got = {(int(a), int(b)): float(v) for a, b, v in zip(m, c, lp)}
    edge = lambda v: abs(v + 10.0) < 1e-3
    assert all(k in got or edge(v) for k, v in ref.items())
    assert all(k in ref or edge(v) for k, v in got.items())
    common = [k for k in ref if k in got]
    err = np.array([abs(ref[k] - got[k]) / max(abs(ref[k]), 1.0) for k in common])
    assert err.max() < 2e-5, err.max()
    # COO order == torch.nonzero order of the reference: sorted by (m, c)
    key = m.astype(np.float64) * 64 + c
    assert np.all(np.diff(key) > 0)
    off_ref = {int(k): int(v) for k, v in zip(m_ref, off_i.numpy().astype(np.int64)) if v != 0 and (int(k),) }
    off_got = dict(zip(off_m.cpu().tolist(), off_v.cpu().tolist()))
    for k, v in off_got.items():
        assert off_ref.get(k) == v
    assert set(off_ref) - set(off_got) <= {k for (k, cc), v in ref.items() if edge(v)}


# ------------------------------------------------------------------------------------------------- BN0
@pytest.mark.parametrize("training", [True, False])
def test_bn0_dropout_matches_torch_batchnorm(ops, training):
    from cellbender_b200 import gemm

    torch.manual_seed(0)
    B, G = 37, 203
    x = ops.alloc_rows(B, G, DEV)
    x[:, :G] = torch.rand(B, G, device=DEV) * 0.01
    bn = torch.nn.BatchNorm1d(G).to(DEV)
    bn.weight.data.uniform_(0.5, 1.5), bn.bias.data.normal_(0, 0.1)
    bn.running_mean.normal_(0, 0.01), bn.running_var.uniform_(0.5, 1.5)
    ref_bn = torch.nn.BatchNorm1d(G).to(DEV).double()
    ref_bn.load_state_dict({k: v.double() if v.is_floating_point() else v for k, v in bn.state_dict().items()})
    bn.train(training), ref_bn.train(training)
    keep = (torch.bernoulli(torch.full((B,), 0.5, device=DEV)) * 2.0) if training else None
    y = gemm.bn0_dropout(x, bn, keep, training, G)
    yr = ref_bn(x[:, :G].double())
    if training:
        yr = yr * keep.double().unsqueeze(-1)
    np.testing.assert_allclose(y.detach().cpu().numpy(), yr.detach().cpu().numpy(), rtol=2e-5, atol=2e-6)
    w = torch.randn(B, G, device=DEV)
    (y * w).sum().backward()
    (yr * w.double()).sum().backward()
    np.testing.assert_allclose(bn.weight.grad.cpu().numpy(), ref_bn.weight.grad.cpu().numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(bn.bias.grad.cpu().numpy(), ref_bn.bias.grad.cpu().numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(bn.running_mean.cpu().numpy(), ref_bn.running_mean.cpu().numpy(), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(bn.running_var.cpu().numpy(), ref_bn.running_var.cpu().numpy(), rtol=1e-5, atol=1e-7)
    assert int(bn.num_batches_tracked) == int(ref_bn.num_batches_tracked)


# ------------------------------------------------------------------------------- simplex parameter transform
@pytest.mark.parametrize("G", [2, 33, 1025, 33538])
def test_simplex_param_matches_torch_transform(ops, G):
    """cb_simplex_forward / _backward vs torch's transform_to(constraints.simplex) (what pyro.param applies to
    chi_ambient, model.py:245-249) evaluated in fp64: value, unit vector and the vector-Jacobian product."""
    from torch.distributions import constraints, transform_to

    from cellbender_b200 import latent

    gen = torch.Generator().manual_seed(G)
    chi0 = torch.distributions.Dirichlet(torch.full((G,), 0.3)).sample().double().clamp_min(1e-12)
    chi0 = chi0 / chi0.sum()
    u32 = (transform_to(constraints.simplex).inv(chi0) + 0.1 * torch.randn(G, generator=gen, dtype=torch.float64)).float()
    u64 = u32.double().requires_grad_(True)
    y64 = transform_to(constraints.simplex)(u64)
    gy = torch.randn(G, generator=gen, dtype=torch.float64)
    (y64 * gy).sum().backward()

    u = u32.to(DEV).requires_grad_(True)
    ld = (G + 3) // 4 * 4
    bufs = {"chi_a": torch.zeros(ld, device=DEV), "chi_unit": torch.zeros(ld, device=DEV)}
    y = latent.simplex_param(u, bufs)
    assert y.shape == (G,) and y.data_ptr() == bufs["chi_a"].data_ptr()
    (y * gy.float().to(DEV)).sum().backward()
    np.testing.assert_allclose(y.detach().cpu().numpy(), y64.detach().numpy(), rtol=2e-6, atol=1e-30)
    assert abs(float(y.sum()) - 1.0) < 1e-5
    unit_ref = y64.detach() / y64.detach().norm()
    np.testing.assert_allclose(bufs["chi_unit"][:G].cpu().numpy(), unit_ref.numpy(), rtol=2e-6, atol=1e-30)
    scale = float(u64.grad.abs().max())
    assert float((u.grad.double().cpu() - u64.grad).abs().max()) <= 1e-5 * scale + 3e-7 * float(gy.abs().max())  # y is stored in fp32
    # without persistent buffers the op allocates its own output
    y2 = latent.simplex_param(u.detach())
    assert torch.equal(y2, y.detach())


# ------------------------------------------------------------------------------- hidden-layer BatchNorm + activation
@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("act", ["softplus", "relu", None])
@pytest.mark.parametrize("B,N", [(255, 512), (7, 40), (64, 33)])
def test_bn_act_matches_torch(ops, training, act, B, N):
    """cb_bn_act_forward / _backward vs nn.BatchNorm1d + activation in fp64 (values, running statistics, input /
    gamma / beta gradients and the column sums of dx handed to the preceding Linear's bias)."""
    from cellbender_b200 import gemm

    gen = torch.Generator().manual_seed(B * 1000 + N)
    x64 = (torch.randn(B, N, generator=gen, dtype=torch.float64) * 3 + 1.5).requires_grad_(True)
    bn64 = torch.nn.BatchNorm1d(N, momentum=0.01, eps=1e-3).double()
    with torch.no_grad():
        bn64.weight.copy_(torch.rand(N, generator=gen, dtype=torch.float64) + 0.5)
        bn64.bias.copy_(torch.randn(N, generator=gen, dtype=torch.float64))
        bn64.running_mean.copy_(torch.randn(N, generator=gen, dtype=torch.float64))
        bn64.running_var.copy_(torch.rand(N, generator=gen, dtype=torch.float64) + 0.5)
    bn = torch.nn.BatchNorm1d(N, momentum=0.01, eps=1e-3).to(DEV)
    bn.load_state_dict({k: v.float() for k, v in bn64.state_dict().items()})
    bn64.train(training), bn.train(training)
    f = {"softplus": torch.nn.functional.softplus, "relu": torch.relu, None: lambda t: t}[act]
    y64 = f(bn64(x64))
    gy = torch.randn(B, N, generator=gen, dtype=torch.float64)
    (y64 * gy).sum().backward()

    lin = torch.nn.Linear(3, N).to(DEV)  # only its bias is used: receives the column sums of dx
    x = x64.detach().float().to(DEV).requires_grad_(True)
    y = gemm.bn_act(x, bn, act, training, bias_of=lin)
    (y * gy.float().to(DEV)).sum().backward()
    tol = dict(rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(y.detach().cpu().numpy(), y64.detach().numpy(), **tol)
    np.testing.assert_allclose(x.grad.cpu().numpy(), x64.grad.numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(bn.weight.grad.cpu().numpy(), bn64.weight.grad.numpy(), rtol=1e-4, atol=2e-4)
    np.testing.assert_allclose(bn.bias.grad.cpu().numpy(), bn64.bias.grad.numpy(), rtol=1e-4, atol=2e-4)
    np.testing.assert_allclose(lin.bias.grad.cpu().numpy(), x64.grad.sum(0).numpy(), rtol=1e-4, atol=3e-4)
    np.testing.assert_allclose(bn.running_mean.cpu().numpy(), bn64.running_mean.numpy(), **tol)
    np.testing.assert_allclose(bn.running_var.cpu().numpy(), bn64.running_var.numpy(), **tol)
    assert int(bn.num_batches_tracked) == int(bn64.num_batches_tracked)


# ------------------------------------------------------------------------------- small per-droplet ops
@pytest.mark.parametrize("B,N,F", [(255, 512, 4), (9, 40, 0), (300, 33, 4)])
def test_head_matches_linear_on_concatenated_input(ops, B, N, F):
    """cb_head_forward / _backward vs nn.Linear(F + N, 1)(cat(feat, x)) in fp64 (EncodeNonZLatents.layer3 and the last
    eps_network layer, vae/encoder.py:216,227,291-293)."""
    gen = torch.Generator().manual_seed(B + N)
    lin64 = torch.nn.Linear(F + N, 1).double()
    x64 = torch.randn(B, N, generator=gen, dtype=torch.float64, requires_grad=True)
    feat64 = torch.randn(B, F, generator=gen, dtype=torch.float64) if F else None
    inp = torch.cat((feat64, x64), dim=-1) if F else x64
    out64 = lin64(inp).squeeze(-1)
    go = torch.randn(B, generator=gen, dtype=torch.float64)
    (out64 * go).sum().backward()

    lin = torch.nn.Linear(F + N, 1).to(DEV)
    lin.load_state_dict({k: v.float() for k, v in lin64.state_dict().items()})
    x = x64.detach().float().to(DEV).requires_grad_(True)
    feat = feat64.float().to(DEV) if F else None
    out = ops.head(x, lin, feat)
    assert out.shape == (B,)
    (out * go.float().to(DEV)).sum().backward()
    np.testing.assert_allclose(out.detach().cpu().numpy(), out64.detach().numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(x.grad.cpu().numpy(), x64.grad.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(lin.weight.grad.cpu().numpy(), lin64.weight.grad.numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(lin.bias.grad.cpu().numpy(), lin64.bias.grad.numpy(), rtol=1e-4, atol=1e-4)


def test_encoder_features_match_reference_formulas(ops):
    """cb_encoder_features vs the feature algebra of EncodeNonZLatents.forward (vae/encoder.py:250-287)."""
    gen = torch.Generator().manual_seed(1)
    B, Z = 77, 64
    feats = torch.stack((torch.randint(0, 50000, (B,), generator=gen).float(), torch.randint(0, 5000, (B,), generator=gen).float(),
                         torch.rand(B, generator=gen) * 100, torch.randint(0, 900, (B,), generator=gen).float()), dim=1)
    z = torch.randn(B, Z, generator=gen)
    d_empty_loc = torch.tensor(4.3)
    log_sum, log_nnz = feats[:, 0].double().log1p(), feats[:, 1].double().log1p()
    overlap = -0.5 * (torch.clamp(log_sum - 4.3, min=0.0) / 0.1) ** 2
    znorm = z.double().norm(dim=-1)
    p_ref = torch.stack((log_sum, log_nnz, overlap, znorm), dim=1)
    e_ref = torch.stack((log_sum, log_nnz, feats[:, 2].double(), znorm), dim=1)
    p, e = ops.encoder_features(feats.to(DEV), z.to(DEV), d_empty_loc.to(DEV))
    np.testing.assert_allclose(p.cpu().numpy(), p_ref.numpy(), rtol=2e-5, atol=1e-5)  # overlap squares a difference / 0.1
    np.testing.assert_allclose(e.cpu().numpy(), e_ref.numpy(), rtol=2e-6, atol=1e-6)
    p0, e0 = ops.encoder_features(feats.to(DEV), z.to(DEV), None)
    assert float(p0[:, 2].abs().max()) == 0.0 and float(e0[:, 2].abs().max()) == 0.0


# ------------------------------------------------------------------------------- exact NB (+) Poisson convolution
def test_nbpc_exact_log_prob_vs_reference_golden_and_float64(ops, golden_dir):
    """cb_nbpc_exact_log_prob_fwd/_bwd (A8: NegativeBinomialPoissonConv.log_prob, consts.USE_EXACT_LOG_PROB) against the
    unmodified reference's output on the golden inputs and float64 autograd of the oracle restatement, with the
    broadcasting the reference uses ([2, B, G] rates, [B, G] counts, scalar alpha)."""
    g = np.load(os.path.join(golden_dir, "nbpc.npz"))
    mask = g["exact_mask"].astype(bool)
    v, mu, al, lam = (torch.tensor(np.ascontiguousarray(g[k][mask], dtype=np.float32), device=DEV) for k in ("value", "mu", "alpha", "lam"))
    out = ops.nbpc_exact_log_prob(v, mu, al, lam)
    ref = g["exact_f64"]
    err = np.abs(out.cpu().numpy() - ref) / np.maximum(np.abs(ref), 1.0)
    small = (v <= 64).cpu().numpy()
    assert err[small].max() < 1e-5, err[small].max()
    assert err.max() < 1e-3, err.max()  # the reference evaluates lgamma of its count tables in float32 (...Conv.py:138-141)

    gen = torch.Generator().manual_seed(0)
    B, G = 6, 40
    x = torch.poisson(torch.rand(B, G, generator=gen) * 6.0, generator=gen)
    mu2 = (torch.rand(2, B, G, generator=gen) * 4 + 0.05).double().requires_grad_(True)
    lam2 = (torch.rand(2, B, G, generator=gen) * 2 + 0.01).double().requires_grad_(True)
    alpha = torch.tensor(3.7, dtype=torch.float64, requires_grad=True)
    w = torch.randn(2, B, G, generator=gen, dtype=torch.float64)
    ref2 = onb.nbpc_exact_log_prob(x.double(), mu2, alpha, lam2)
    (ref2 * w).sum().backward()
    mu_c, lam_c = mu2.detach().float().to(DEV).requires_grad_(True), lam2.detach().float().to(DEV).requires_grad_(True)
    al_c = alpha.detach().float().to(DEV).requires_grad_(True)
    got = ops.nbpc_exact_log_prob(x.to(DEV), mu_c, al_c, lam_c)
    assert got.shape == (2, B, G)
    (got * w.float().to(DEV)).sum().backward()
    np.testing.assert_allclose(got.detach().cpu().numpy(), ref2.detach().numpy(), rtol=1e-5, atol=1e-5)
    for mine, r in ((mu_c.grad, mu2.grad), (lam_c.grad, lam2.grad), (al_c.grad, alpha.grad)):
        scale = float(r.abs().max())
        assert float((mine.double().cpu() - r).abs().max()) <= 1e-4 * scale


# ------------------------------------------------------------------------------- PRq posterior regularisation
@pytest.mark.parametrize("tag", ["fixture", "rand0", "rand1", "rand2"])
def test_prq_matches_reference_golden(ops, golden_dir, tag):
    """regularization.PRq (cb_prq_regularize) vs PRq.regularize / _compute_log_target_dict of the unmodified reference
    (posterior.py:1002-1225): same kept entries, regularised log probs to 1e-6, targets to 1e-9, for alpha 0, 1, 2."""
    from cellbender_b200.regularization import PRq

    g = np.load(os.path.join(golden_dir, "prq.npz"))
    shape = tuple(int(v) for v in g[f"{tag}_shape"])
    coo = sp.coo_matrix((g[f"{tag}_data"], (g[f"{tag}_row"], g[f"{tag}_col"])), shape=shape)
    offsets = dict(zip(g[f"{tag}_off_m"].tolist(), g[f"{tag}_off_v"].tolist()))
    for alpha in (0.0, 1.0, 2.0):
        key = f"{tag}_a{alpha:g}"
        tgt = PRq._compute_log_target_dict(noise_count_posterior_coo=coo, alpha=alpha)
        ref_t = dict(zip(g[key + "_target_m"].tolist(), g[key + "_target"].tolist()))
        assert set(tgt) == set(ref_t)
        for k_, v in ref_t.items():
            assert (np.isneginf(v) and np.isneginf(tgt[k_])) or abs(v - tgt[k_]) < 1e-6, (k_, v, tgt[k_])
        reg = PRq.regularize(noise_count_posterior_coo=coo, noise_offsets=offsets, alpha=alpha, device=DEV, target_tolerance=0.001)
        assert reg.shape == shape and reg.data.dtype == np.float64
        np.testing.assert_array_equal(reg.row, g[key + "_row"])
        np.testing.assert_array_equal(reg.col, g[key + "_col"])
        np.testing.assert_allclose(reg.data, g[key + "_data"], rtol=1e-6, atol=1e-6)


def test_prq_large_random_vs_oracle_and_mean_targets(ops):
    """A larger posterior (thousands of segments, offsets, holes): kernel vs the numpy oracle entry by entry, and the
    reference test's own criterion (tests/test_posterior.py:78-137): regularised means hit mean + alpha * std."""
    from cellbender_b200.estimation import DevicePosteriorCOO
    from cellbender_b200.regularization import PRq

    rng = np.random.default_rng(5)
    rows, cols, vals, offsets = [], [], [], {}
    n_m, n_c = 3000, 20
    for m in range(n_m):
        k = int(rng.integers(1, n_c + 1))
        lp = rng.normal(size=k) * 2.0
        lp = lp - np.log(np.exp(lp).sum())
        keep = lp >= -7
        keep[np.argmax(lp)] = True
        for c in np.flatnonzero(keep):
            rows.append(m * 7 + 3), cols.append(c), vals.append(lp[c])
        if rng.random() < 0.3:
            offsets[m * 7 + 3] = int(rng.integers(1, 5))
    coo = sp.coo_matrix((np.array(vals, np.float32), (np.array(rows), np.array(cols))), shape=(n_m * 7 + 10, n_c))
    alpha = 1.0
    m_o, c_o, lp_o, _ = opost.prq_regularize(coo.row, coo.col, coo.data, offsets, alpha)
    reg = PRq.regularize(noise_count_posterior_coo=coo, noise_offsets=offsets, alpha=alpha, device=DEV)
    np.testing.assert_array_equal(reg.row, m_o)
    np.testing.assert_array_equal(reg.col, c_o)
    np.testing.assert_allclose(reg.data, lp_o, rtol=1e-6, atol=1e-6)
    # device-to-device variant keeps (m, c) order and feeds the estimators
    dev_reg = PRq.regularize_device(DevicePosteriorCOO.from_scipy(coo, offsets, DEV), alpha=alpha)
    assert torch.equal(dev_reg.m.cpu(), torch.as_tensor(m_o)) and dev_reg.log_prob.dtype == torch.float32
    # means: E_reg[n] == E[n] + alpha std[n] wherever that target is reachable (n = c + offset enters the tilt only)
    dense = np.full((n_m, n_c), -np.inf)
    dense[(coo.row - 3) // 7, coo.col] = coo.data
    cnt = np.arange(n_c)[None, :]
    p_in = np.exp(dense)
    mean_in = (p_in * cnt).sum(1)
    std_in = np.sqrt((p_in * (cnt - mean_in[:, None]) ** 2).sum(1))
    dreg = np.full((n_m, n_c), -np.inf)
    dreg[(reg.row - 3) // 7, reg.col] = reg.data
    off_arr = np.array([offsets.get(m * 7 + 3, 0) for m in range(n_m)])
    mean_reg_n = (np.exp(dreg) * (cnt + off_arr[:, None])).sum(1)  # the constraint is on E[c + offset]
    target = mean_in + alpha * std_in
    multi = (np.isfinite(dense).sum(1) > 1) & (target > 0)
    reachable = multi & (target < (np.where(np.isfinite(dense), cnt + off_arr[:, None], -1).max(1) - 1e-3)) \
        & (target > (np.where(np.isfinite(dense), cnt + off_arr[:, None], 10 ** 6).min(1) + 1e-3))
    assert reachable.sum() > 500
    np.testing.assert_allclose(mean_reg_n[reachable], target[reachable], rtol=2e-3, atol=2e-3)


def test_torch_binary_search_matches_reference_golden(golden_dir):
    from cellbender_b200.regularization import torch_binary_search

    g = np.load(os.path.join(golden_dir, "prq.npz"))
    tgt = torch.tensor(g["bsearch_target"], device=DEV)
    out = torch_binary_search(evaluate_outcome_given_value=lambda x: x ** 2, target_outcome=tgt,
                              init_range=torch.tensor([0.0, 100.0], device=DEV).unsqueeze(0).expand(4, 2),
                              target_tolerance=0.001, max_iterations=100)
    np.testing.assert_array_equal(out.cpu().numpy(), g["bsearch_out"])
    with pytest.raises(AssertionError):
        torch_binary_search(lambda x: x, tgt, torch.zeros(4, 2, device=DEV), target_tolerance=0.0)


@pytest.mark.parametrize("tag", ["case0", "case1"])
@pytest.mark.parametrize("per_gene", [False, True])
@pytest.mark.parametrize("fpr", [0.01, 0.3])
def test_prmu_matches_reference_golden(ops, golden_dir, tag, per_gene, fpr):
    """regularization.PRmu (cb_posterior_tilt + device MAP + bisection) vs PRmu.regularize of the unmodified reference
    (posterior.py:1228-1530), overall and per-gene factors: same kept entries, log probs to 1e-6."""
    from cellbender_b200.estimation import IndexConverter
    from cellbender_b200.regularization import PRmu

    g = np.load(os.path.join(golden_dir, "prmu.npz"))
    shape = tuple(int(v) for v in g[f"{tag}_shape"])
    nc, ng = (int(v) for v in g[f"{tag}_ncng"])
    coo = sp.coo_matrix((g[f"{tag}_data"], (g[f"{tag}_row"], g[f"{tag}_col"])), shape=shape)
    offsets = dict(zip(g[f"{tag}_off_m"].tolist(), g[f"{tag}_off_v"].tolist()))
    raw = sp.csr_matrix(g[f"{tag}_raw"])
    np.random.seed(0)
    reg = PRmu.regularize(noise_count_posterior_coo=coo, noise_offsets=offsets, index_converter=IndexConverter(nc, ng),
                          raw_count_matrix=raw, fpr=fpr, per_gene=per_gene, device=DEV, n_cells=1000)
    key = f"{tag}_{'gene' if per_gene else 'all'}_f{fpr:g}"
    assert reg.shape == shape
    np.testing.assert_array_equal(reg.row, g[key + "_row"])
    np.testing.assert_array_equal(reg.col, g[key + "_col"])
    np.testing.assert_allclose(reg.data, g[key + "_data"], rtol=1e-6, atol=1e-6)
    with pytest.raises(AssertionError):
        PRmu.regularize(noise_count_posterior_coo=coo, noise_offsets=offsets, raw_count_matrix=raw)


# ---------------------------------------------------------------------------------------------- fused P2P Adam
@pytest.mark.parametrize("n_ranks,own", [(1, 0), (2, 0), (2, 1), (3, 1), (4, 3), (8, 5), (16, 0)])
def test_clipped_adam_p2p_matches_sum_then_sweep(ops, n_ranks, own):
    """cb_clipped_adam_p2p on ONE device (the 'peer' buffers are local allocations): gradients of all ranks summed in
    rank order + ClippedAdam on the shard + parameter stores into every peer buffer must equal sum -> cb_clipped_adam_range
    -> copy (TMA bulk prefetch of the peers' gradient tiles); ragged last tile; step counter advanced."""
    import ctypes

    from cellbender_b200 import _cabi

    g = torch.Generator(device=DEV).manual_seed(n_ranks * 17 + own)
    n_total, a = 4 * (256 * 37 + 13) + 64, 64  # the shard starts 64 elements into the buffers
    n = n_total - a
    rnd = lambda scale=1.0: torch.randn(n_total, device=DEV, generator=g) * scale
    p0, m0, v0 = rnd(), rnd(0.1), rnd(0.1).abs()
    grads = [rnd(8.0) for _ in range(n_ranks)]  # some values beyond the +-10 clamp
    lr = torch.tensor([1e-3, 2e-3, 3e-3], device=DEV)
    b1 = torch.tensor([0.95, 0.9, 0.85], device=DEV)
    L = _cabi.lib()
    st = _cabi.current_stream()
    # reference: rank-order sum, then the single-GPU sweep
    gsum = grads[0].clone()
    for x in grads[1:]:
        gsum += x
    p_ref, m_ref, v_ref = p0.clone(), m0.clone(), v0.clone()
    step_ref = torch.tensor([1], dtype=torch.int64, device=DEV)
    at = lambda t: t.data_ptr() + 4 * a
    L.call("cb_clipped_adam_range", at(p_ref), at(gsum), at(m_ref), at(v_ref), n, lr.data_ptr(), b1.data_ptr(), None, 3,
           step_ref.data_ptr(), 0.999, 1e-8, 10.0, 1, st)
    # fused kernel
    p, m, v = p0.clone(), m0.clone(), v0.clone()
    peers = [torch.full((n_total,), 7.0, device=DEV) for _ in range(n_ranks - 1)]
    step = torch.tensor([1], dtype=torch.int64, device=DEV)
    g_arr = (ctypes.c_void_p * n_ranks)(*[at(x) for x in grads])
    p_arr = (ctypes.c_void_p * max(1, n_ranks - 1))(*[at(x) for x in peers])
    L.call("cb_clipped_adam_p2p", at(p), at(m), at(v), n, ctypes.addressof(g_arr), n_ranks, own, ctypes.addressof(p_arr),
           n_ranks - 1, lr.data_ptr(), b1.data_ptr(), 3, step.data_ptr(), 0.999, 1e-8, 10.0, 1, st)
    torch.cuda.synchronize()
    assert int(step) == 2 and int(step_ref) == 2
    for got, want in ((p, p_ref), (m, m_ref), (v, v_ref)):
        assert torch.equal(got[:a], want[:a]), "elements before the shard were touched"
        torch.testing.assert_close(got[a:], want[a:], rtol=1e-6, atol=1e-8)
    assert float((p[a:] - p0[a:]).abs().max()) > 0
    for peer in peers:
        assert torch.equal(peer[a:], p[a:]), "peer parameter buffer must receive exactly the owner's new values"
        assert float(peer[:a].sub(7.0).abs().max()) == 0.0
    with pytest.raises(_cabi.CabiError):
        L.call("cb_clipped_adam_p2p", at(p), at(m), at(v), n - 1, ctypes.addressof(g_arr), n_ranks, own, ctypes.addressof(p_arr),
               n_ranks - 1, lr.data_ptr(), b1.data_ptr(), 3, step.data_ptr(), 0.999, 1e-8, 10.0, 1, st)
