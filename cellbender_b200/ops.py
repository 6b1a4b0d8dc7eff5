"""Torch-facing wrappers of the C-ABI kernels.  torch is plumbing here (device memory, streams);
all compute is in ``libcellbender_b200.so``.  Every function requires CUDA tensors and raises if the
extension is missing -- there is no CPU path in the product."""
from dataclasses import dataclass
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _cabi
from ._cabi import current_stream, ptr
from .exceptions import NanException


def round_up(n: int, k: int) -> int:
    return (n + k - 1) // k * k


ROW_ALIGN = 32  # floats


def row_pitch(n_cols: int) -> int:
    """Leading dimension of a dense fp32 matrix with ``n_cols`` logical columns: a multiple of 32 floats, so that every row
    starts on a 128-byte line.  (A pitch that is only 16-byte aligned -- round_up(33 538, 4) = 33 540 -- makes every
    128-byte TMA box row and every warp-wide row segment straddle two lines: 5 sectors moved for 4 used, twice the L2
    requests.  The wide [B, G] kernels measured 2.4-3.4 TB/s with that pitch, the contiguous optimizer sweep 6 TB/s.)"""
    return round_up(n_cols, ROW_ALIGN)


def _req_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("cellbender_b200 ops require CUDA tensors (no CPU fallback)")


_SIDE_STREAMS = {}


LOW_PRIORITY_STREAMS = (0, 4)


def side_stream(device, idx: int = 0) -> "torch.cuda.Stream":
    """A persistent auxiliary stream per (device, idx).  Independent kernels of one step (the weight-gradient and
    input-gradient products of a layer; torch's slow reparameterisation-gradient kernels) are forked onto it and joined
    again, which under CUDA-graph capture becomes parallel branches of the step graph.  Callers keep every tensor that
    crosses streams alive until the join, so the caching allocator never recycles one early."""
    d = torch.device(device)
    key = (d.index if d.index is not None else torch.cuda.current_device(), idx)
    if key not in _SIDE_STREAMS:
        # Stream 0 carries the weight-gradient products and the optimizer sweeps behind them, stream 4 the column sums of
        # the observation kernel's chi_ambient gradient (consumed at the very end of backward): bulk HBM traffic with no
        # consumer on the critical chain.  Every other stream carries links of the latency-bound critical chain and gets
        # the higher priority, so that its (small) kernels are placed ahead of the bulk CTAs still waiting for a slot.
        # Priorities are recorded into the kernel nodes of a captured CUDA graph.
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=key[0], priority=0 if idx in LOW_PRIORITY_STREAMS else -1)
    return _SIDE_STREAMS[key]


def alloc_rows(n_rows: int, n_cols: int, device, zero: bool = True) -> torch.Tensor:
    """Row-major [n_rows, ld] fp32 with ld = row_pitch(n_cols); use ``[:, :n_cols]`` for the logical matrix."""
    ld = row_pitch(n_cols)
    f = torch.zeros if zero else torch.empty
    return f((n_rows, ld), dtype=torch.float32, device=device)


@dataclass
class DeviceCSR:
    """Count matrix resident in HBM: indptr int64[R+1], indices int32[nnz] (sorted per row), data fp32[nnz]."""

    indptr: torch.Tensor
    indices: torch.Tensor
    data: torch.Tensor
    shape: Tuple[int, int]

    @staticmethod
    def from_scipy(mat, device) -> "DeviceCSR":
        import scipy.sparse as sp

        mat = sp.csr_matrix(mat)
        mat.sum_duplicates()
        mat.sort_indices()
        return DeviceCSR(
            torch.as_tensor(mat.indptr.astype(np.int64), device=device),
            torch.as_tensor(mat.indices.astype(np.int32), device=device),
            torch.as_tensor(mat.data.astype(np.float32), device=device),
            (int(mat.shape[0]), int(mat.shape[1])),
        )

    @property
    def n_genes(self) -> int:
        return self.shape[1]


def gather_rows(csr: DeviceCSR, row_ids: torch.Tensor, amb_unit: Optional[torch.Tensor] = None,
                want: Sequence[str] = ("raw", "norm", "lognorm"), out: Optional[Dict[str, torch.Tensor]] = None):
    """cb_csr_gather_rows.  Returns dict with the requested dense [B, ld] matrices and 'feats' [B, 4]
    (sum, nnz, dot(x, amb_unit), max)."""
    _req_cuda(csr.indptr, row_ids)
    B, G = int(row_ids.numel()), csr.n_genes
    dev = row_ids.device
    out = {} if out is None else out
    for k in want:
        if k not in out:
            out[k] = alloc_rows(B, G, dev, zero=False)
    if "feats" not in out:
        out["feats"] = torch.empty((B, 4), dtype=torch.float32, device=dev)
    lds = {int(out[k].stride(0)) for k in want}
    assert len(lds) <= 1, "the dense outputs of one gather must share a leading dimension"
    ld = lds.pop() if lds else row_pitch(G)
    _cabi.lib().call(
        "cb_csr_gather_rows", ptr(csr.indptr), ptr(csr.indices), ptr(csr.data), ptr(row_ids), B, G, ld, ptr(amb_unit),
        ptr(out.get("raw") if "raw" in want else None), ptr(out.get("norm") if "norm" in want else None),
        ptr(out.get("lognorm") if "lognorm" in want else None), ptr(out["feats"]), current_stream())
    return out


# ------------------------------------------------------------------------------------------------------
# NBPCapprox.log_prob as a differentiable op (drop-in for the reference distribution's log_prob)
# ------------------------------------------------------------------------------------------------------
def _strides3(t: torch.Tensor, shape3):
    """view `t` broadcast to the 3-D logical shape; broadcast dims get element stride 0"""
    t3 = t.reshape((1,) * (3 - t.dim()) + tuple(t.shape)).expand(shape3)
    return t3, tuple(t3.stride())


def _as3(*ts):
    shape = torch.broadcast_shapes(*[t.shape for t in ts])
    if len(shape) > 3:
        lead = int(np.prod(shape[:-2]))
        shape3 = (lead, shape[-2], shape[-1])
        ts = [t.expand(shape).reshape(shape3) for t in ts]  # may copy for exotic broadcasts
    else:
        shape3 = (1,) * (3 - len(shape)) + tuple(shape)
    return shape, shape3, ts


class _NbpcApproxLogProb(torch.autograd.Function):
    """log_prob of NegativeBinomialPoissonConvApprox (max_poisson < 0) or of the exact convolution
    NegativeBinomialPoissonConv (max_poisson >= 0: number of Poisson terms - 1)."""

    @staticmethod
    def forward(ctx, value, mu, alpha, lam, max_poisson=-1):
        _req_cuda(value, mu, alpha, lam)
        args = [t.float() for t in (value, mu, alpha, lam)]
        shape, shape3, args = _as3(*args)
        packed = [_strides3(t, shape3) for t in args]
        import ctypes

        sarr = [(ctypes.c_longlong * 3)(*s) for _, s in packed]
        out = torch.empty(shape3, dtype=torch.float32, device=value.device)
        flag = torch.zeros(1, dtype=torch.int32, device=value.device)
        E, B, G = shape3
        operands = (ptr(packed[0][0]), ctypes.addressof(sarr[0]), ptr(packed[1][0]), ctypes.addressof(sarr[1]),
                    ptr(packed[2][0]), ctypes.addressof(sarr[2]), ptr(packed[3][0]), ctypes.addressof(sarr[3]))
        if max_poisson < 0:
            _cabi.lib().call("cb_nbpc_approx_log_prob_fwd", *operands, ptr(out), E, B, G, ptr(flag), current_stream())
        else:
            _cabi.lib().call("cb_nbpc_exact_log_prob_fwd", *operands, ptr(out), E, B, G, int(max_poisson), ptr(flag),
                             current_stream())
        ctx.max_poisson = max_poisson
        ctx.save_for_backward(*[p[0] for p in packed])
        ctx.strides = [p[1] for p in packed]
        ctx.shape, ctx.shape3 = shape, shape3
        ctx.in_shapes = [t.shape for t in (value, mu, alpha, lam)]
        ctx.nan_flag = flag
        return out.reshape(shape)

    @staticmethod
    def backward(ctx, grad_out):
        import ctypes

        v3, m3, a3, l3 = ctx.saved_tensors
        sarr = [(ctypes.c_longlong * 3)(*s) for s in ctx.strides]
        E, B, G = ctx.shape3
        go = grad_out.reshape(ctx.shape3).contiguous().float()
        dmu, dlam, dal = (torch.empty(ctx.shape3, dtype=torch.float32, device=go.device) for _ in range(3))
        operands = (ptr(v3), ctypes.addressof(sarr[0]), ptr(m3), ctypes.addressof(sarr[1]), ptr(a3), ctypes.addressof(sarr[2]),
                    ptr(l3), ctypes.addressof(sarr[3]), ptr(go), ptr(dmu), ptr(dlam), ptr(dal), E, B, G)
        if ctx.max_poisson < 0:
            _cabi.lib().call("cb_nbpc_approx_log_prob_bwd", *operands, current_stream())
        else:
            _cabi.lib().call("cb_nbpc_exact_log_prob_bwd", *operands, int(ctx.max_poisson), current_stream())

        def red(g, shp):
            return g.reshape(ctx.shape).sum_to_size(shp) if tuple(shp) != tuple(ctx.shape) else g.reshape(ctx.shape)

        return None, red(dmu, ctx.in_shapes[1]), red(dal, ctx.in_shapes[2]), red(dlam, ctx.in_shapes[3]), None


def nbpc_approx_log_prob(value, mu, alpha, lam, check_nan: bool = True):
    """log_prob of NegativeBinomialPoissonConvApprox(mu, alpha, lam) at `value`, broadcasting like the
    reference (...ConvApprox.py:105-143).  Raises NanException like the reference when a NaN appears."""
    alpha = torch.as_tensor(alpha, dtype=torch.float32, device=mu.device)
    out = _NbpcApproxLogProb.apply(value, mu, alpha, lam)
    if check_nan and bool(torch.isnan(out.sum())):
        bad = [n for n, t in (("mu", mu), ("alpha", alpha), ("lam", lam)) if bool(torch.isnan(t.log().sum()))]
        raise NanException(param=", ".join(bad))
    return out


def nbpc_exact_log_prob(value, mu, alpha, lam, max_poisson: int = 100, check_nan: bool = True):
    """log_prob of NegativeBinomialPoissonConv(mu, alpha, lam, max_poisson) at `value`: the exact NB (+) Poisson
    convolution (reference distributions/NegativeBinomialPoissonConv.py:89-154), differentiable in mu, alpha, lam."""
    alpha = torch.as_tensor(alpha, dtype=torch.float32, device=mu.device)
    out = _NbpcApproxLogProb.apply(value, mu, alpha, lam, int(max_poisson))
    if check_nan and bool(torch.isnan(out.sum())):
        bad = [n for n, t in (("mu", mu), ("alpha", alpha), ("lam", lam)) if bool(torch.isnan(t.log().sum()))]
        raise NanException(param=", ".join(bad))
    return out


# ------------------------------------------------------------------------------------------------------
# fused ELBO observation terms
# ------------------------------------------------------------------------------------------------------
SUM_NAMES = ("L1", "L0", "LE", "dL1_da_mu1", "dL1_dcA1", "dL1_dcB1", "dL1_dalpha", "dL0_dcA0", "dL0_dcB0",
             "dL0_dalpha", "dLE_dcAE", "dLE_dcBE")


def elbo_obs_scratch_elems(n_rows: int) -> int:
    """fp32 elements of scratch cb_elbo_obs_fwd_bwd needs for ``n_rows`` droplets."""
    return int(_cabi.lib().raw("cb_elbo_obs_scratch_elems")(int(n_rows)))


def elbo_obs(csr: DeviceCSR, row_ids, logits, chi_ambient, chi_bar, coef, alpha, w, *, out=None, defer_colsum: bool = False,
             max_poisson: Optional[int] = None):
    """cb_elbo_obs_fwd_bwd (+ cb_colsum_rows).  logits: [B, ld] buffer (ld % 4 == 0) holding G valid columns.
    chi_ambient / chi_bar: fp32 [>= G] 16-byte aligned.  ``max_poisson``: None = the approximate likelihood (the
    reference's default), an int = the exact convolution with that many Poisson terms (cb_elbo_obs_exact_fwd_bwd).
    Returns dict(sums [B,12], dlogits [B,ld], dchi_ambient [G], lse [B], nan_flag)."""
    _req_cuda(logits, chi_ambient, chi_bar, coef, alpha, w, row_ids)
    B, G = int(row_ids.numel()), csr.n_genes
    ld = logits.stride(0)
    assert logits.shape[0] == B and ld % 4 == 0 and ld >= G and logits.stride(1) == 1
    dev = logits.device
    out = {} if out is None else out
    if "sums" not in out:
        out["sums"] = torch.empty((B, 12), dtype=torch.float32, device=dev)
    if "dlogits" not in out:
        out["dlogits"] = torch.zeros((B, ld), dtype=torch.float32, device=dev)
    if "qamb" not in out:
        out["qamb"] = torch.empty((B, ld), dtype=torch.float32, device=dev)
    if "dchi_ambient" not in out:
        out["dchi_ambient"] = torch.empty(G, dtype=torch.float32, device=dev)
    if "lse" not in out:
        out["lse"] = torch.empty(B, dtype=torch.float32, device=dev)
    if "nan_flag" not in out:
        out["nan_flag"] = torch.zeros(1, dtype=torch.int32, device=dev)
    st = current_stream()
    L = _cabi.lib()
    n_scratch = int(L.raw("cb_elbo_obs_scratch_elems")(B))
    if "scratch" not in out or out["scratch"].numel() < n_scratch:
        out["scratch"] = torch.empty(n_scratch, dtype=torch.float32, device=dev)  # per-tile partial sums (float64) + lse pairs
    operands = (ptr(csr.indptr), ptr(csr.indices), ptr(csr.data), ptr(row_ids), B, G, ptr(logits), ld, ptr(chi_ambient), ptr(chi_bar),
                ptr(coef), ptr(alpha), ptr(w), ptr(out["sums"]), ptr(out["dlogits"]), ptr(out["qamb"]), ld, ptr(out["lse"]),
                ptr(out["nan_flag"]), ptr(out["scratch"]))
    if max_poisson is None:
        L.call("cb_elbo_obs_fwd_bwd", *operands, st)
    else:
        L.call("cb_elbo_obs_exact_fwd_bwd", *operands, int(max_poisson), st)
    if not defer_colsum:
        L.call("cb_colsum_rows", ptr(out["qamb"]), B, G, ld, ptr(out["dchi_ambient"]), st)
        out["dchi_ready"] = None
        return out
    # training step: the column sums (d loss / d chi_ambient) are consumed only at the very end of backward (simplex
    # VJP), so they are issued on an auxiliary stream; out["dchi_ready"] is the event the consumer must wait for
    main = torch.cuda.current_stream()
    side = side_stream(logits.device, 4)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        L.call("cb_colsum_rows", ptr(out["qamb"]), B, G, ld, ptr(out["dchi_ambient"]), current_stream())
        ev = torch.cuda.Event()
        ev.record(side)
    out["dchi_ready"] = ev
    return out


def colsum_rows(mat: torch.Tensor, n_cols: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _req_cuda(mat)
    out = torch.empty(n_cols, dtype=torch.float32, device=mat.device) if out is None else out
    _cabi.lib().call("cb_colsum_rows", ptr(mat), mat.shape[0], n_cols, mat.stride(0), ptr(out), current_stream())
    return out


def col_logsumexp(mat: torch.Tensor, n_rows: int, n_cols: int, row_add: Optional[torch.Tensor] = None,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[c] = log sum_r exp(mat[r, c] + row_add[r]) over the first n_rows x n_cols block of ``mat`` (cb_col_logsumexp)."""
    _req_cuda(mat, row_add)
    L = _cabi.lib()
    out = torch.empty(n_cols, dtype=torch.float32, device=mat.device) if out is None else out
    scratch = torch.empty(int(L.raw("cb_col_logsumexp_scratch_elems")(n_rows, n_cols)), dtype=torch.float32, device=mat.device)
    L.call("cb_col_logsumexp", ptr(mat), n_rows, n_cols, mat.stride(0), ptr(row_add), ptr(scratch), ptr(out), current_stream())
    return out


def clipped_adam_step(p, g, m, v, lr_table, beta1_table, step, beta2=0.999, eps=1e-8, clip=10.0, grad_scale=1.0,
                      step_table=None):
    """cb_clipped_adam_step over flat fp32 buffers; `step` is a device int64[1] counter (incremented).  ``step_table``:
    optional per-step bias-corrected step sizes (optim.FlatClippedAdam builds them in float64 on the host)."""
    _req_cuda(p, g, m, v, lr_table, beta1_table, step, step_table)
    _cabi.lib().call("cb_clipped_adam_step", ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), ptr(lr_table), ptr(beta1_table),
                     ptr(step_table), lr_table.numel(), ptr(step), float(beta2), float(eps), float(clip), float(grad_scale),
                     current_stream())


def exclusive_scan_i32(x: torch.Tensor):
    """Exclusive prefix sum int32 -> int64 positions; returns (positions, total as 0-dim device tensor)."""
    _req_cuda(x)
    n = x.numel()
    L = _cabi.lib()
    out = torch.empty(n, dtype=torch.int64, device=x.device)
    scratch = torch.empty(int(L.raw("cb_scan_scratch_elems")(n)), dtype=torch.int64, device=x.device)
    total = torch.zeros(1, dtype=torch.int64, device=x.device)
    L.call("cb_exclusive_scan_i32", ptr(x), n, ptr(out), ptr(scratch), ptr(total), current_stream())
    return out, total


# ------------------------------------------------------------------------------------------------------
# posterior noise-count COO
# ------------------------------------------------------------------------------------------------------
class PosteriorWorkspace:
    """Per-stored-count outputs of the noise-window kernel for a whole posterior pass (all chunks), sized once from the
    cells' stored-count numbers so that no host synchronisation is needed between chunks: logp [E, stride], low [E],
    keep [E], entry_m [E]."""

    def __init__(self, n_entries: int, n_counts_max: int, device):
        self.n_entries = int(n_entries)  # rows in use (<= capacity)
        self.capacity = max(int(n_entries), 1)
        self.n_counts_max = int(n_counts_max)
        self.stride = int(_cabi.lib().raw("cb_posterior_window_stride")(int(n_counts_max)))
        n = max(self.n_entries, 1)
        self.logp = torch.empty((n, self.stride), dtype=torch.float32, device=device)
        self.low = torch.empty(n, dtype=torch.int32, device=device)
        self.keep = torch.empty(n, dtype=torch.int32, device=device)
        self.entry_m = torch.empty(n, dtype=torch.int64, device=device)


@torch.no_grad()
def posterior_window(csr: DeviceCSR, cell_rows, logits_t, dec_bias, lse, chi_ambient, chi_bar, a_mu, c_a, c_b, alpha, n_window,
                     entry_start, n_entries: int, barcode_all, gene_all, total_n_genes: int, ws: PosteriorWorkspace,
                     entry_offset, log_thresh: float = -10.0, lambda_multiplier: float = 1.0, ctas_per_sm: int = 0):
    """cb_posterior_noise_window for one chunk of cells, writing rows [entry_offset, entry_offset + entry_start[-1]) of
    ``ws``.  ``entry_offset``: host int or a device int64 scalar tensor; ``n_entries``: the chunk's stored counts or an
    upper bound of them (it only sizes the grid) -- both forms exist so that one captured graph can serve all chunks.

    logits_t [G, ldt]: transposed, cell-major decoder output (column b * S + s); lse [Bc * S]; a_mu, c_a, c_b [Bc * S];
    alpha [S]; n_window [Bc] int32; entry_start [Bc + 1] int64 (chunk-local prefix sums of the stored-count numbers);
    barcode_all [Bc] / gene_all [G] int64 map to indices of the full count matrix.  ``ctas_per_sm`` > 0 caps the grid."""
    _req_cuda(cell_rows, logits_t, lse, a_mu, c_a, c_b, alpha, n_window, entry_start, barcode_all, gene_all)
    Bc, G, S = int(cell_rows.numel()), csr.n_genes, int(alpha.numel())
    assert logits_t.stride(1) == 1 and logits_t.shape[0] >= G
    if isinstance(entry_offset, torch.Tensor):
        assert entry_offset.dtype == torch.int64 and entry_offset.is_cuda
        off_ptr, o = ptr(entry_offset), 0
    else:
        off_ptr, o = None, int(entry_offset)
        assert o + n_entries <= max(ws.n_entries, 1)
    _cabi.lib().call(
        "cb_posterior_noise_window", ptr(csr.indptr), ptr(csr.indices), ptr(csr.data), ptr(cell_rows), Bc, G, S, ptr(logits_t),
        logits_t.stride(0), ptr(dec_bias), ptr(lse), ptr(chi_ambient), ptr(chi_bar), ptr(a_mu), ptr(c_a), ptr(c_b), ptr(alpha),
        ptr(n_window), ptr(entry_start), int(n_entries), ptr(barcode_all), ptr(gene_all), int(total_n_genes), ws.n_counts_max,
        float(log_thresh), float(lambda_multiplier), off_ptr, ws.logp.data_ptr() + 4 * ws.stride * o, ws.low.data_ptr() + 4 * o,
        ws.keep.data_ptr() + 4 * o, ws.entry_m.data_ptr() + 8 * o, int(ctas_per_sm), current_stream())


@torch.no_grad()
def posterior_compact(ws: PosteriorWorkspace, log_thresh: float = -10.0):
    """Thresholded COO of a filled workspace: (m int64, c int32, log_prob fp32, off_m int64, off_v int32), entries in
    workspace order (cell, gene, c).  The ONE host synchronisation of a posterior pass (two output sizes) happens here."""
    L = _cabi.lib()
    st = current_stream()
    dev = ws.logp.device
    n = ws.n_entries
    if n == 0:
        z = lambda dt: torch.zeros(0, dtype=dt, device=dev)
        return z(torch.int64), z(torch.int32), z(torch.float32), z(torch.int64), z(torch.int32)
    keep = ws.keep[:n]
    keep_pos, keep_total = exclusive_scan_i32(keep)
    has_off = torch.empty(n, dtype=torch.int32, device=dev)
    L.call("cb_posterior_offset_flags", n, ptr(keep), ptr(ws.low), ptr(has_off), st)
    off_pos, off_total = exclusive_scan_i32(has_off)
    n_keep, n_off = (int(v) for v in torch.cat((keep_total, off_total)).tolist())
    coo_m = torch.empty(n_keep, dtype=torch.int64, device=dev)
    coo_c = torch.empty(n_keep, dtype=torch.int32, device=dev)
    coo_lp = torch.empty(n_keep, dtype=torch.float32, device=dev)
    off_m = torch.empty(n_off, dtype=torch.int64, device=dev)
    off_v = torch.empty(n_off, dtype=torch.int32, device=dev)
    L.call("cb_posterior_compact", n, ws.n_counts_max, ptr(ws.entry_m), ptr(ws.logp), ptr(ws.low), ptr(keep), ptr(keep_pos),
           ptr(off_pos), float(log_thresh), ptr(coo_m), ptr(coo_c), ptr(coo_lp), ptr(off_m), ptr(off_v), st)
    return coo_m, coo_c, coo_lp, off_m, off_v


@torch.no_grad()
def posterior_coo(csr: DeviceCSR, cell_rows, logits_t, dec_bias, lse, chi_ambient, chi_bar, a_mu, c_a, c_b, alpha, n_window,
                  barcode_all, gene_all, total_n_genes: int, n_counts_max: int = 20, log_thresh: float = -10.0,
                  lambda_multiplier: float = 1.0):
    """One-chunk convenience wrapper (tests): window kernel + compaction.  See ``posterior_window`` for the operands."""
    dev = logits_t.device
    nnz_per_cell = (csr.indptr[cell_rows + 1] - csr.indptr[cell_rows])
    entry_start = torch.zeros(int(cell_rows.numel()) + 1, dtype=torch.int64, device=dev)
    entry_start[1:] = torch.cumsum(nnz_per_cell, 0)
    n_entries = int(entry_start[-1].item())
    ws = PosteriorWorkspace(n_entries, n_counts_max, dev)
    posterior_window(csr, cell_rows, logits_t, dec_bias, lse, chi_ambient, chi_bar, a_mu, c_a, c_b, alpha, n_window, entry_start,
                     n_entries, barcode_all, gene_all, total_n_genes, ws, 0, log_thresh, lambda_multiplier)
    return posterior_compact(ws, log_thresh)


def gather_feats(csr: DeviceCSR, row_ids: torch.Tensor, amb_unit: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Per-droplet features only ([B, 4]: sum, nnz, dot with amb_unit, max) -- no dense rows are written."""
    return gather_rows(csr, row_ids, amb_unit, want=())["feats"]


# ------------------------------------------------------------------------------------------------------
# tcgen05 TF32 GEMM
# ------------------------------------------------------------------------------------------------------
def gemm_tf32(a: torch.Tensor, b: torch.Tensor, M: int, N: int, K: int, a_mn: bool = False, b_mn: bool = False,
              out: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None, accumulate: bool = False,
              split_k: Optional[int] = None, a2: Optional[torch.Tensor] = None, b2: Optional[torch.Tensor] = None,
              K2: int = 0, tile_cfg: int = 0) -> torch.Tensor:
    """C[M,N] (=,+=) A[M,K] @ B[N,K]^T (+ bias) through cb_gemm_tf32.

    a: stored [M, >=K] (a_mn=False) or [K, >=M] (a_mn=True), last dim contiguous, row stride % 4 == 0; same for b."""
    _req_cuda(a, b)
    L = _cabi.lib()
    assert a.stride(-1) == 1 and b.stride(-1) == 1 and a.dtype == torch.float32 and b.dtype == torch.float32
    dev = a.device
    if out is None:
        out = torch.empty((M, row_pitch(N)), dtype=torch.float32, device=dev)[:, :N]
    if split_k is None:
        split_k = int(L.raw("cb_gemm_tf32_suggest_split")(M, N, K + K2))
    ws_elems = int(L.raw("cb_gemm_tf32_workspace_elems")(M, N, split_k))
    # split-K partial tiles: a fresh allocation per call from torch's stream-ordered caching allocator.  A captured CUDA
    # graph refers to the workspace by address, so it must never be a cached buffer that a later call (another minibatch
    # size, the posterior pass) could replace and free: the allocator keeps a captured graph's memory in that graph's pool.
    ws = torch.empty(ws_elems, dtype=torch.float32, device=dev) if ws_elems else None
    if tile_cfg:  # explicit tile configuration (tools/gemm_bench.py)
        a2_, b2_ = (a, b) if a2 is None else (a2, b2)
        L.call("cb_gemm_tf32_ex", ptr(a), a.stride(0), ptr(b), b.stride(0), K, ptr(a2_), a2_.stride(0), ptr(b2_), b2_.stride(0), K2,
               int(a_mn), int(b_mn), ptr(out), out.stride(0), M, N, ptr(bias), int(accumulate), split_k, ptr(ws), ws_elems,
               int(tile_cfg), current_stream())
    elif a2 is None:
        L.call("cb_gemm_tf32", ptr(a), a.stride(0), int(a_mn), ptr(b), b.stride(0), int(b_mn), ptr(out), out.stride(0), M, N, K,
               ptr(bias), int(accumulate), split_k, ptr(ws), ws_elems, current_stream())
    else:  # C = A B^T + A2 B2^T in one launch (shared TMEM accumulator)
        assert a2.stride(-1) == 1 and b2.stride(-1) == 1 and K2 > 0
        L.call("cb_gemm_tf32_pair", ptr(a), a.stride(0), ptr(b), b.stride(0), K, ptr(a2), a2.stride(0), ptr(b2), b2.stride(0), K2,
               int(a_mn), int(b_mn), ptr(out), out.stride(0), M, N, ptr(bias), int(accumulate), split_k, ptr(ws), ws_elems,
               current_stream())
    return out


# ------------------------------------------------------------------------------------------------------
# small per-droplet ops (csrc/small_ops.cu)
# ------------------------------------------------------------------------------------------------------
@torch.no_grad()
def encoder_features(feats: torch.Tensor, z: torch.Tensor, d_empty_loc: Optional[torch.Tensor]):
    """(p_feat, e_feat) [B, 4] of EncodeNonZLatents.forward (vae/encoder.py:250-287); ``d_empty_loc`` = constrained
    pyro.param value (0-dim / 1-element tensor) or None when no ambient profile is given.  Inputs are data: no gradient."""
    _req_cuda(feats, z, d_empty_loc)
    B = feats.shape[0]
    z = z.detach().reshape(B, -1).contiguous()
    p_feat = torch.empty((B, 4), dtype=torch.float32, device=feats.device)
    e_feat = torch.empty((B, 4), dtype=torch.float32, device=feats.device)
    _cabi.lib().call("cb_encoder_features", ptr(feats.contiguous()), ptr(z), z.stride(0), B, z.shape[1],
                     ptr(d_empty_loc.detach().float().contiguous()) if d_empty_loc is not None else None, ptr(p_feat), ptr(e_feat),
                     current_stream())
    return p_feat, e_feat


class _Head(torch.autograd.Function):
    """Linear(F + N -> 1)(cat(feat, x)).squeeze(-1) through cb_head_forward / _backward; weight and bias gradients go
    straight into the optimizer's flat gradient buffer when the parameters carry a view of it."""

    @staticmethod
    def forward(ctx, x, weight, bias, feat):
        B, N = x.shape
        if x.stride(1) != 1:
            x = x.contiguous()
        F_ = 0 if feat is None else feat.shape[1]
        out = torch.empty(B, dtype=torch.float32, device=x.device)
        _cabi.lib().call("cb_head_forward", ptr(x), x.stride(0), ptr(feat), F_, B, N, ptr(weight), ptr(bias), ptr(out),
                         current_stream())
        ctx.save_for_backward(x, weight) if feat is None else ctx.save_for_backward(x, weight, feat)
        ctx.params = (weight, bias)
        return out

    @staticmethod
    def backward(ctx, d_out):
        from .gemm import _grad_target

        saved = ctx.saved_tensors
        x, weight = saved[0], saved[1]
        feat = saved[2] if len(saved) > 2 else None
        w_param, b_param = ctx.params
        B, N = x.shape
        F_ = 0 if feat is None else feat.shape[1]
        dx = torch.empty((B, N), dtype=torch.float32, device=x.device) if ctx.needs_input_grad[0] else None
        dw_buf, dw = _grad_target(w_param)
        db_buf, db = (None, None)
        if b_param is not None and ctx.needs_input_grad[2]:
            db_buf, db = _grad_target(b_param)
        _cabi.lib().call("cb_head_backward", ptr(x), x.stride(0), ptr(feat), F_, B, N, ptr(weight), ptr(d_out.contiguous()),
                         ptr(dx), dx.stride(0) if dx is not None else 0, ptr(dw_buf), ptr(db_buf), current_stream())
        return dx, dw, db, None


def head(x: torch.Tensor, lin: torch.nn.Linear, feat: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``lin(cat(feat, x)).squeeze(-1)`` for a one-output nn.Linear whose first feat.shape[1] input columns are features."""
    _req_cuda(x, feat)
    assert lin.out_features == 1 and lin.weight.is_contiguous()
    return _Head.apply(x, lin.weight, lin.bias, None if feat is None else feat.contiguous())


def feat_dw(dy: torch.Tensor, feat: torch.Tensor, dw: torch.Tensor):
    """dw[:, :F] = dy^T @ feat (cb_feat_dw): the feature columns of the weight gradient of Linear(cat(feat, x))."""
    _req_cuda(dy, feat, dw)
    B, H = dy.shape
    F_ = feat.shape[1]
    assert dy.stride(1) == 1 and feat.is_contiguous() and dw.stride(1) == 1 and dw.shape[0] == H
    _cabi.lib().call("cb_feat_dw", ptr(dy), dy.stride(0), ptr(feat), F_, B, H, ptr(dw), dw.stride(0), current_stream())


def obs_loss(sums: torch.Tensor, w: torch.Tensor):
    """(sum_b w . L, its negative) as two 0-dim tensors (cb_obs_loss)."""
    B = sums.shape[0]
    loss = torch.empty((), dtype=torch.float32, device=sums.device)
    neg = torch.empty((), dtype=torch.float32, device=sums.device)
    _cabi.lib().call("cb_obs_loss", ptr(sums), sums.stride(0), ptr(w), B, ptr(loss), ptr(neg), current_stream())
    return loss, neg


def obs_small_grads(sums: torch.Tensor, w: torch.Tensor, g: Optional[torch.Tensor], want_dw: bool = True):
    """(d_coef [B, 7], d_alpha [1], d_w [B, 3] or None) of loss = g * sum_b w . L (cb_obs_small_grads)."""
    B = sums.shape[0]
    f32 = dict(dtype=torch.float32, device=sums.device)
    d_coef, d_alpha = torch.empty((B, 7), **f32), torch.empty(1, **f32)
    d_w = torch.empty((B, 3), **f32) if want_dw else None
    _cabi.lib().call("cb_obs_small_grads", ptr(sums), sums.stride(0), ptr(w), ptr(g), B, ptr(d_coef), ptr(d_alpha), ptr(d_w),
                     current_stream())
    return d_coef, d_alpha, d_w
