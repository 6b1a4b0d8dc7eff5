"""Numerics of the kernels' fp32 device math, run on the CPU: the math headers are __host__ __device__, so the
very same source is compiled for the host (tools/host_math_check.cu) and compared with the multiprecision /
float64 truth.  This is how kernel numerics are iterated on in the GPU-less build container; the -m gpu tests
repeat the comparison through the real kernels."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import posterior as opost

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hostmath") / "hostmath.so")
    subprocess.run(["nvcc", "-O2", "-shared", "-Xcompiler", "-fPIC", "-Wno-deprecated-gpu-targets", "-o", so,
                    os.path.join(ROOT, "tools", "host_math_check.cu")], check=True)
    return ctypes.CDLL(so)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_nbpc_fp32_math_vs_multiprecision(hostlib, golden_dir):
    g = np.load(os.path.join(golden_dir, "nbpc.npz"))
    v, mu, al, lam = (np.ascontiguousarray(g[k], dtype=np.float32) for k in ("value", "mu", "alpha", "lam"))
    n = v.size
    L, dmu, dlam, dal = (np.zeros(n, np.float32) for _ in range(4))
    hostlib.host_nbpc_eval(n, _p(v), _p(mu), _p(al), _p(lam), _p(L), _p(dmu), _p(dlam), _p(dal))
    ref = g["approx_mp"]
    err = np.abs(L - ref) / np.maximum(np.abs(ref), 1.0)
    # tolerance of north_star: 1e-5 relative (to max(|L|,1)) for every count (counts above 64 take the float64 pipe)
    assert err.max() < 1e-5, (err.max(), v[np.argmax(err)])
    # the reference's own fp32 evaluation is far worse on the same inputs (context for the tolerance)
    ref32 = np.abs(g["approx_f32"] - ref) / np.maximum(np.abs(ref), 1.0)
    assert np.nanmax(ref32) > 1e-2
    for name, mine in (("dmu", dmu), ("dlam", dlam), ("dalpha", dal)):
        r = g[f"approx_{name}_mp"]
        e = np.abs(mine - r) / np.maximum(np.abs(r), 1e-3)
        assert e.max() < 1e-3, (name, e.max())
        assert np.mean(e > 1e-4) < 0.01, (name, np.mean(e > 1e-4))


def test_nbpc_zero_count_specialisation_vs_multiprecision(hostlib):
    """nbpc_zero (the x == 0 closed form the dense pass of the fused observation kernel evaluates 2 x 33 538 times
    per droplet) against an 80-digit evaluation of the reference formula and of its analytic derivatives, over the
    rate ranges the model produces (mu << lam up to mu >> lam), including the Poisson branch mu < 1e-5 lam."""
    from oracle import nbpc as onb

    rng = np.random.default_rng(3)
    n = 4000
    mu = np.exp(rng.uniform(np.log(1e-10), np.log(50.0), n)).astype(np.float32)
    lam = np.exp(rng.uniform(np.log(1e-10), np.log(5.0), n)).astype(np.float32)
    mu[:50] = 1e-10  # the y = 0 branch: mu is the bare safeguard constant
    alpha = np.exp(rng.uniform(np.log(0.05), np.log(500.0), n)).astype(np.float32)
    v = np.zeros(n, np.float32)
    L, dmu, dlam, dal = (np.zeros(n, np.float32) for _ in range(4))
    hostlib.host_nbpc_zero(n, _p(mu), _p(alpha), _p(lam), _p(L), _p(dmu), _p(dlam), _p(dal))
    Lg, dmug, dlamg, dalg = (np.zeros(n, np.float32) for _ in range(4))
    hostlib.host_nbpc_eval(n, _p(v), _p(mu), _p(alpha), _p(lam), _p(Lg), _p(dmug), _p(dlamg), _p(dalg))
    ref_L, ref_dmu, ref_dlam, ref_dal = onb.nbpc_approx_mp(v.astype(np.float64), mu.astype(np.float64),
                                                           alpha.astype(np.float64), lam.astype(np.float64))
    poisson = mu < np.float32(1e-5) * lam
    # Poisson branch: the reference overwrites L and passes no gradient to mu / alpha
    ref_L = np.where(poisson, -lam.astype(np.float64), ref_L)
    ref_dmu, ref_dal = np.where(poisson, 0.0, ref_dmu), np.where(poisson, 0.0, ref_dal)
    ref_dlam = np.where(poisson, -1.0, ref_dlam)
    assert np.max(np.abs(L - ref_L) / np.maximum(np.abs(ref_L), 1.0)) < 2e-6
    for name, mine, generic, r in (("dmu", dmu, dmug, ref_dmu), ("dlam", dlam, dlamg, ref_dlam), ("dalpha", dal, dalg, ref_dal)):
        scale = np.maximum(np.abs(r), 1e-3)
        assert np.max(np.abs(mine - r) / scale) < 1e-4, name
        assert np.max(np.abs(mine - generic) / scale) < 1e-4, name  # and it agrees with the general-count path

    # the y = 0 collapse: mu is exactly the 1e-10 safeguard; lam spans the Poisson-branch boundary (1e-5 lam vs 1e-10)
    n0 = 3000
    lam0 = np.exp(rng.uniform(np.log(1e-10), np.log(5.0), n0)).astype(np.float32)
    al0 = np.exp(rng.uniform(np.log(0.05), np.log(500.0), n0)).astype(np.float32)
    mu0 = np.full(n0, 1e-10, np.float32)
    L0, dl0, da0 = (np.zeros(n0, np.float32) for _ in range(3))
    hostlib.host_nbpc_zero_mu_eps(n0, _p(al0), _p(lam0), _p(L0), _p(dl0), _p(da0))
    rL, _, rdl, rda = onb.nbpc_approx_mp(np.zeros(n0), mu0.astype(np.float64), al0.astype(np.float64), lam0.astype(np.float64))
    assert np.max(np.abs(L0 - rL) / np.maximum(np.abs(rL), 1e-30)) < 2e-7  # relative, down to lam = 1e-10
    assert np.max(np.abs(dl0 - rdl)) < 1e-6
    assert np.max(np.abs(da0 - rda) / np.maximum(np.abs(rda), 1e-30)) < 1e-5  # values ~1e-20 .. 1e-18: relative


@pytest.mark.parametrize("tag", ["big", "small", "huge"])
def test_noise_window_fp32_math_vs_float64(hostlib, golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "posterior_logprob.npz"))
    data, mu, lam, alpha = (g[f"{tag}_{k}"] for k in ("data", "mu", "lam", "alpha"))
    d32, m32, l32, a32 = (np.asarray(a, np.float32) for a in (data, mu + 1e-30, lam + 1e-30, alpha + 1e-30))
    lp, low = opost.log_prob_noise_count_tensor(
        torch.tensor(d32, dtype=torch.float64), torch.tensor(m32, dtype=torch.float64),
        torch.tensor(l32, dtype=torch.float64), torch.tensor(float(a32), dtype=torch.float64), n_counts_max=20,
        mimic_float_casts=False)
    lpn = (lp - torch.logsumexp(lp, -1, keepdim=True)).numpy()
    n = lp.shape[-1]
    nz = np.nonzero(d32 > 0)
    x, m_, l_ = (np.ascontiguousarray(a[nz]) for a in (d32, m32, l32))
    a_ = np.full(x.size, a32, np.float32)
    prob = np.zeros((x.size, 32), np.float32)
    lo = np.zeros(x.size, np.int32)
    hostlib.host_noise_window_pmf(x.size, _p(x), _p(m_), _p(l_), _p(a_), n, _p(prob), _p(lo))
    np.testing.assert_array_equal(lo, low.numpy()[nz].astype(np.int32))
    ref = lpn[nz]
    with np.errstate(divide="ignore"):
        mine = np.log(prob[:, :n])
    keep = ref >= -10.0  # the posterior keeps log prob >= -10 (posterior.py:426,533)
    err = np.abs(mine - ref)[keep] / np.maximum(np.abs(ref[keep]), 1.0)
    assert err.max() < 1e-5, err.max()


def test_nbpc_exact_convolution_vs_float64_and_reference_golden(hostlib, golden_dir):
    """nbpc_exact_eval (the kernel math of cb_nbpc_exact_log_prob_fwd/_bwd, run on the host) against (a) a float64
    evaluation of NegativeBinomialPoissonConv.log_prob's definition (scipy gammaln / logsumexp over the same 101-term
    window), (b) the unmodified reference's own output on the golden inputs (which evaluates its count tables' lgamma
    in float32, ...Conv.py:138-141, hence the looser bound for large counts), (c) float64 autograd of the oracle
    restatement for the three gradients."""
    from scipy.special import gammaln, logsumexp

    from oracle import nbpc as onb

    g = np.load(os.path.join(golden_dir, "nbpc.npz"))
    mask = g["exact_mask"].astype(bool)
    v, mu, al, lam = (np.ascontiguousarray(g[k][mask], dtype=np.float32) for k in ("value", "mu", "alpha", "lam"))
    n = v.size
    assert n > 200 and (v == 0).any() and (v == 1).any() and (v > 100).any()
    L, dmu, dlam, dal = (np.zeros(n, np.float32) for _ in range(4))
    hostlib.host_nbpc_exact(n, _p(v), _p(mu), _p(al), _p(lam), 100, _p(L), _p(dmu), _p(dlam), _p(dal))

    v64, mu64, al64, lam64 = (a.astype(np.float64) for a in (v, mu, al, lam))
    low = np.clip(np.minimum((lam - np.float32(50.0)).astype(np.int32), (v - np.float32(100.0)).astype(np.int32)), 0, None)
    k = np.arange(101)[None, :]
    pois = low[:, None] + k
    nb = v64[:, None] - pois
    ok = nb >= 0
    nbc = np.where(ok, nb, 0.0)
    a_, m_, l_ = al64[:, None], mu64[:, None], lam64[:, None]
    terms = (pois * np.log(l_) - l_ - gammaln(pois + 1.0)
             + gammaln(nbc + a_) - gammaln(nbc + 1.0) - gammaln(a_) + a_ * (np.log(a_) - np.log(a_ + m_))
             + nbc * (np.log(m_) - np.log(a_ + m_)))
    ref = logsumexp(np.where(ok, terms, -np.inf), axis=1)
    err = np.abs(L - ref) / np.maximum(np.abs(ref), 1.0)
    assert err.max() < 2e-6, err.max()

    ref_golden = g["exact_f64"]  # already restricted to exact_mask: the reference itself (float32 lgamma of the count tables inside)
    err_g = np.abs(L - ref_golden) / np.maximum(np.abs(ref_golden), 1.0)
    assert err_g[v <= 64].max() < 1e-5, err_g[v <= 64].max()
    assert err_g.max() < 1e-3, err_g.max()

    # gradients: softmax-weighted term gradients in float64 (analytic), all counts
    from scipy.special import digamma

    wgt = np.where(ok, np.exp(terms - ref[:, None]), 0.0)
    g_lam = (wgt * (pois / l_ - 1.0)).sum(1)
    g_mu = (wgt * (nbc / m_ - (a_ + nbc) / (a_ + m_))).sum(1)
    g_al = (wgt * (digamma(nbc + a_) - digamma(a_) + np.log(a_ / (a_ + m_)) + 1.0 - (a_ + nbc) / (a_ + m_))).sum(1)
    for name, mine, r in (("dmu", dmu, g_mu), ("dlam", dlam, g_lam), ("dalpha", dal, g_al)):
        e = np.abs(mine - r) / np.maximum(np.abs(r), 1e-3)
        assert e.max() < 1e-4, (name, e.max())  # fp32 logs inside the 101-term pmf recurrences
    # and the oracle restatement's float64 autograd agrees where its float32 count-table lgamma is harmless
    t = lambda a: torch.tensor(a, dtype=torch.float64, requires_grad=True)
    mu_t, al_t, lam_t = t(mu64), t(al64), t(lam64)
    onb.nbpc_exact_log_prob(torch.tensor(v64), mu_t, al_t, lam_t).sum().backward()
    small = v <= 16
    for name, mine, r in (("dmu", dmu, mu_t.grad), ("dlam", dlam, lam_t.grad), ("dalpha", dal, al_t.grad)):
        r = r.numpy()
        e = np.abs(mine - r) / np.maximum(np.abs(r), 1e-3)
        assert e[small].max() < 1e-3, (name, e[small].max())
