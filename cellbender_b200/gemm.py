"""The large contractions of a training step (three G -> 512 encoder input layers, one 512 -> G decoder output
layer; reference vae/encoder.py:97-105,211,220 and vae/decoder.py:41-47), their backward, and the gene-wise
BatchNorm + row dropout in front of them (vae/encoder.py:239,279).

Every product goes through csrc/gemm_tf32.cu -- hand-written TMA-fed tcgen05.mma kind::tf32 with TMEM accumulators
and a TMA-store epilogue.  TF32 multiply (10-bit mantissa, truncation), fp32 accumulation.  Stated tolerance vs an
fp32/fp64 product: 2e-3 of the output scale per element (tests/test_gemm_gpu.py); on the ELBO: <= 2e-3 relative
(tests/test_step_gpu.py::test_elbo_tf32_tolerance); end to end: cell calls and removed counts within 1 % of the
reference-style fp32 pipeline (tests/test_e2e_gpu.py).  There is no other back-end: operands that TMA cannot address
(unaligned base, row pitch not a multiple of 16 bytes) are an error, not a reason to fall back to a library GEMM.  The
parity tests that need full-fp32 products substitute ``ops.gemm_tf32`` from inside tests/ (tests/conftest.py).
Weight gradients are written by the GEMM epilogue straight into the optimizer's flat gradient buffer (the
nn.Parameter carries its gradient view as ``_cb_grad_view``), so no extra copy or accumulation pass exists.
"""
from typing import Optional

import torch

from . import ops
from ._cabi import current_stream, lib, ptr

# optimizer whose piecewise sweep is issued behind the weight-gradient GEMMs during the current backward pass (set by
# train.SVI around loss.backward(); None = gradients are only written to the flat gradient buffer)
FUSED_OPT = None
# only matrices with at least this many elements are swept early (= optim.BIG_BLOCK_NUMEL: the blocks the optimizer lays
# out first)
FUSED_MIN_NUMEL = 1 << 22


def _grad_target(weight: torch.Tensor):
    """(buffer the weight gradient is written to, what backward returns to autograd).  When the optimizer has given
    the parameter a view of its flat gradient buffer, the GEMM writes there and autograd is handed None (so that
    AccumulateGrad neither clones nor re-accumulates 68 MB); ``_cb_grad_direct`` tells ``collect_grads`` about it."""
    gv = getattr(weight, "_cb_grad_view", None)
    if gv is None or not weight.requires_grad:
        t = torch.empty_like(weight)
        return t, t
    weight._cb_grad_direct = True
    return gv, None


# the persistent 0-dim tensor 1.0 that train.SVI backpropagates from (``loss.backward(UNIT_GRAD)``): no ones_like fill per
# step, and backward functions that recognise it by address skip their multiplications by the upstream gradient
UNIT_GRAD = None

_PENDING_JOINS = []
_BULK_PENDING = []
_CALLBACK_QUEUED = []  # non-empty while the running backward pass has its end-of-backward join callback queued
BULK_STREAM, SMALL_GRAD_STREAM, LATE_GRAD_STREAM = 0, 5, 6  # ops.side_stream indices


def grad_stream(device, weights) -> "torch.cuda.Stream":
    """Stream for the weight-gradient products of a layer: the bulk stream (low priority; also carries the optimizer sweeps
    of the large blocks) for the G-wide matrices, a high-priority one for the small layers -- queued behind ~300 us of
    bulk sweeps their 8 us products used to finish last and hold up the end of the step (r02h timeline)."""
    big = any(w.numel() >= FUSED_MIN_NUMEL for w in weights)
    return ops.side_stream(device, BULK_STREAM if big else SMALL_GRAD_STREAM)


def late_grad_stream(device) -> "torch.cuda.Stream":
    """High-priority stream for the G-wide weight-gradient products of the encoders' input layers.  They run at the very end
    of backward, when nothing of the critical chain is left to protect: on the low-priority bulk stream they queued behind
    the optimizer sweeps of the blocks finished before them and crawled (r02y timeline: 201 us for a 37 us product), which
    postponed their own sweeps -- the tail of the step.  Only the sweeps stay on the bulk stream."""
    return ops.side_stream(device, LATE_GRAD_STREAM)


def defer_join(side: "torch.cuda.Stream"):
    """Postpone ``current_stream().wait_stream(side)`` to the end of the running backward pass.  Weight / bias gradients
    that a forked kernel writes straight into the optimizer's flat gradient buffer have no consumer inside backward, so
    the calling stream (which carries the critical chain of input-gradient kernels) need not stall for them at the end
    of the node; the autograd engine's final callback joins every such stream before control returns to the caller --
    except the bulk stream, which the optimizer joins as late as it can (``join_bulk``: after the sweep of everything
    the bulk stream does not touch)."""
    if not _CALLBACK_QUEUED:
        def _join_all():
            _CALLBACK_QUEUED.clear()
            main = torch.cuda.current_stream()
            while _PENDING_JOINS:
                main.wait_stream(_PENDING_JOINS.pop())
        _CALLBACK_QUEUED.append(True)
        torch.autograd.Variable._execution_engine.queue_callback(_join_all)
    bulk = ops.side_stream(side.device, BULK_STREAM)
    if side is bulk and FUSED_OPT is not None:  # an optimizer step follows and will join (train.SVI sets FUSED_OPT)
        if side not in _BULK_PENDING:
            _BULK_PENDING.append(side)
    elif side not in _PENDING_JOINS:
        _PENDING_JOINS.append(side)


def join_bulk():
    """Make the current stream wait for the bulk stream(s) forked during the last backward pass (no-op when none)."""
    main = torch.cuda.current_stream()
    while _BULK_PENDING:
        main.wait_stream(_BULK_PENDING.pop())


def early_update_applies(w: torch.Tensor) -> bool:
    opt = FUSED_OPT
    return (opt is not None and getattr(opt, "early_weight_updates", False) and getattr(w, "_cb_block", None) is not None
            and getattr(w, "_cb_grad_view", None) is not None and w.requires_grad and w.numel() >= FUSED_MIN_NUMEL)


def _tma_addressable(t: torch.Tensor, col_offset: int = 0) -> bool:
    return (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.stride(1) == 1 and t.stride(0) % 4 == 0
            and (t.data_ptr() + 4 * col_offset) % 16 == 0)


def _require_tma(t: torch.Tensor, what: str, col_offset: int = 0):
    """The tensor-core path is the only path: fail loudly when an operand cannot be addressed by TMA."""
    if not _tma_addressable(t, col_offset):
        raise RuntimeError(
            f"cellbender_b200.gemm: {what} (shape {tuple(t.shape)}, strides {tuple(t.stride())}, dtype {t.dtype}, device "
            f"{t.device}) is not TMA-addressable: needs a CUDA fp32 matrix with unit column stride, a row pitch that is a "
            f"multiple of 4 floats and a 16-byte aligned base.  Large weights get such a layout from "
            f"gemm.pad_big_weights(module) / optim.FlatClippedAdam; there is no library-GEMM fallback.")


def needs_padded_layout(p: torch.Tensor) -> bool:
    """2-D weights that get a padded storage block: the large ones (128-byte-aligned rows) and every one whose row length
    is not a multiple of 4 floats (TMA needs 16-byte rows: Linear(6 -> 512) with --z-dim 6, Linear(516 -> 512) is fine)."""
    return p.dim() == 2 and (p.numel() >= (1 << 16) or p.shape[1] % 4 != 0)


def pad_big_weights(module: torch.nn.Module):
    """Re-home every 2-D weight of ``module`` that needs it (``needs_padded_layout``) into storage TMA can address; names,
    shapes and values are unchanged (the nn.Parameter becomes a [rows, cols] view of a [rows, ld] block whose padding
    columns stay zero).  The optimizer does the same when it builds its flat buffer (optim.FlatClippedAdam); this covers
    models that never see an optimizer (posterior from a checkpoint)."""
    with torch.no_grad():
        for p in module.parameters():
            if needs_padded_layout(p):
                lead, ld = padded_layout(p)
                if p.stride(1) == 1 and p.stride(0) == ld and (p.data_ptr() - 4 * lead) % (128 if ld % 32 == 0 else 16) == 0:
                    continue
                rows, cols = p.shape
                block = torch.zeros((rows, ld), dtype=p.dtype, device=p.device)
                block[:, lead:lead + cols].copy_(p.data)
                p.data = block[:, lead:lead + cols]
    return module


def padded_layout(p: torch.Tensor):
    """(lead, ld) of the storage block of a large 2-D weight: ``ld`` floats per row (a multiple of 32: rows start on
    128-byte lines) of which the first ``lead`` are padding.  ``p._cb_lead`` (set by the module that owns the weight)
    shifts the row so that the G-wide block of a Linear(cat(feat, x)) weight -- which starts ``col_offset`` columns into
    the row -- begins on a 128-byte line too."""
    lead = int(getattr(p, "_cb_lead", 0))
    if p.numel() < (1 << 16) and lead == 0:
        return 0, ops.round_up(p.shape[1], 4)  # small layer: 16-byte rows are all TMA needs
    return lead, ops.row_pitch(lead + p.shape[1])


class LazyFeat:
    """Place-holder for the hand-made feature block of Linear(cat(feat, x)) when the gene-wide product is launched before
    the features exist (they need EncodeZ's output): the forward product then covers the gene block only, ``bn_act`` adds
    the feature term, and the backward pass (which runs long after ``value`` has been set) computes the feature columns
    of the weight gradient as usual."""

    def __init__(self):
        self.value: Optional[torch.Tensor] = None


def _stacked_rows(w0: torch.Tensor, w1: torch.Tensor) -> bool:
    """True when w1's rows directly follow w0's in memory with the same row pitch (consecutive blocks of the optimizer's
    flat buffer): [w0; w1] then is ONE matrix."""
    return (w0.dim() == 2 and w1.dim() == 2 and w0.shape[1] == w1.shape[1] and w0.stride() == w1.stride() and w0.stride(1) == 1
            and w0.dtype == w1.dtype and w1.data_ptr() == w0.data_ptr() + w0.element_size() * w0.shape[0] * w0.stride(0))


class _MultiLinearBig(torch.autograd.Function):
    """Several nn.Linear layers that read the SAME wide input x (EncodeNonZLatents' layer1 and eps_network.0, or the
    single EncodeZ input layer):
         y_i = x[:, :n_in] @ W_i[:, off:off+n_in]^T + feat_i @ W_i[:, :off]^T + bias_i
    One autograd node, so the input gradient dx = sum_i dy_i @ W_i is accumulated by the GEMM epilogue (no separate
    add pass over [B, G]) and every dW_i lands directly in the optimizer's flat gradient buffer."""

    @staticmethod
    def forward(ctx, x, n_in: int, col_offset: int, streams, *args):
        feats, weights, biases = args[0::3], args[1::3], args[2::3]
        B = x.shape[0]
        ys = []
        main = torch.cuda.current_stream()
        fused = (len(weights) == 2 and all(isinstance(f, LazyFeat) for f in feats) and _stacked_rows(weights[0], weights[1])
                 and (biases[0] is None) == (biases[1] is None))
        if fused:
            # The two towers' input layers sit in consecutive blocks of the flat parameter buffer: [W_0; W_1] is one
            # [H_0 + H_1, ld] matrix, so both products are ONE launch that reads the wide input once (two separate deep-ring
            # launches cannot share an SM -- 200 KB of shared memory each -- and ran one after the other: r02z timeline)
            w0, w1 = weights
            H0, H1 = w0.shape[0], w1.shape[0]
            wcat = torch.as_strided(w0, (H0 + H1, w0.shape[1]), w0.stride())
            bias = None if biases[0] is None else torch.cat([biases[0].reshape(-1), biases[1].reshape(-1)])
            y = torch.empty((B, ops.row_pitch(H0 + H1)), dtype=torch.float32, device=x.device)
            ops.gemm_tf32(x, wcat[:, col_offset:col_offset + n_in], B, H0 + H1, n_in, bias=bias, out=y[:, :H0 + H1])
            ys = [y[:, :H0], y[:, H0:H0 + H1]]
            if streams is not None:
                done = torch.cuda.Event()
                done.record(main)
                for st in streams:
                    if st is not None:  # the caller consumes that output on this stream
                        st.wait_event(done)
                        y.record_stream(st)
        for i, (feat, w, b) in enumerate(zip(feats, weights, biases) if not fused else ()):
            st = streams[i] if streams is not None else None
            if st is not None:  # this product runs on its own stream (the caller joins it before consuming the output)
                y = torch.empty((B, ops.row_pitch(w.shape[0])), dtype=torch.float32, device=x.device)[:, :w.shape[0]]
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    assert feat is None or isinstance(feat, LazyFeat)
                    ops.gemm_tf32(x, w[:, col_offset:col_offset + n_in], B, w.shape[0], n_in, bias=b, out=y)
                y.record_stream(st), x.record_stream(st)
            elif isinstance(feat, LazyFeat):
                y = ops.gemm_tf32(x, w[:, col_offset:col_offset + n_in], B, w.shape[0], n_in, bias=b)
            elif feat is not None:
                # Linear(cat(feat, x)): the few feature columns ride along as a second operand pair (one extra k-block)
                if not (col_offset % 4 == 0 and feat.is_contiguous() and feat.shape[1] == col_offset):
                    raise RuntimeError("multi_linear_big: the feature block must be contiguous [B, col_offset] with col_offset % 4 == 0")
                y = ops.gemm_tf32(x, w[:, col_offset:col_offset + n_in], B, w.shape[0], n_in, bias=b,
                                  a2=feat, b2=w[:, :col_offset], K2=col_offset)
            else:
                y = ops.gemm_tf32(x, w[:, col_offset:col_offset + n_in], B, w.shape[0], n_in, bias=b)
            ys.append(y)
        ctx.lazy = [f if isinstance(f, LazyFeat) else None for f in feats]
        feats = [None if isinstance(f, LazyFeat) else f for f in feats]
        ctx.save_for_backward(x, *[t for t in feats if t is not None], *weights)
        ctx.meta = (len(weights), n_in, col_offset, [f is not None for f in feats], [b is not None for b in biases])
        ctx.biases = biases  # the nn.Parameter objects (they carry the optimizer's gradient views)
        ctx.weights_p = weights
        return tuple(ys)

    @staticmethod
    def backward(ctx, *dys):
        n, n_in, off, has_feat, has_bias = ctx.meta
        saved = list(ctx.saved_tensors)
        x = saved.pop(0)
        feats = [saved.pop(0) if h else None for h in has_feat]
        feats = [lz.value if lz is not None else f for f, lz in zip(feats, ctx.lazy)]
        weights = saved
        weights_p = ctx.weights_p
        B = x.shape[0]
        need_dx = ctx.needs_input_grad[0]
        dys = [d.contiguous() for d in dys]
        dx = ops.alloc_rows(B, n_in, x.device, zero=False)[:, :n_in] if need_dx else None
        # every output buffer is allocated here, on the calling stream, before any fork
        targets = []
        for i in range(n):
            w = weights[i]
            dw_t = _grad_target(weights_p[i]) if ctx.needs_input_grad[4 + 3 * i + 1] else None
            db_t = _grad_target(ctx.biases[i]) if (has_bias[i] and ctx.needs_input_grad[4 + 3 * i + 2]) else None
            targets.append((dw_t, db_t))

        early = [i for i in range(n) if targets[i][0] is not None and early_update_applies(weights_p[i])
                 and targets[i][0][1] is None]
        dw_done = {}

        def parameter_gradients():
            for i in range(n):
                dy, w, feat = dys[i], weights[i], feats[i]
                H = w.shape[0]
                dw_t, db_t = targets[i]
                if dw_t is not None:
                    # dW[H, n_in] = dy^T[H, B] . x[B, n_in]: both operands stored with the reduction dim (B) as rows
                    ops.gemm_tf32(dy, x, H, n_in, B, a_mn=True, b_mn=True, out=dw_t[0][:, off:off + n_in], split_k=1)
                    if off:  # the few feature columns in front of the gene block
                        if feat is None:
                            raise RuntimeError("multi_linear_big: a Linear with feature columns was differentiated without its features")
                        ops.feat_dw(dy, feat, dw_t[0])
                if db_t is not None:
                    ops.colsum_rows(dy, H, out=db_t[0])
                if i in early:
                    dw_done[i] = torch.cuda.Event()
                    dw_done[i].record()

        # the weight-gradient products and the input-gradient product are independent: fork the former onto a side
        # stream (parallel branches of the captured step graph), join before returning
        big = any(w.numel() >= FUSED_MIN_NUMEL for w in weights)
        fork = (need_dx or big) and any(t[0] is not None or t[1] is not None for t in targets)
        main = torch.cuda.current_stream()
        if fork:
            side = late_grad_stream(x.device) if big else grad_stream(x.device, weights)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                parameter_gradients()
        else:
            parameter_gradients()
        if need_dx:
            if n == 2:
                # dx = dy_0 . W_0 + dy_1 . W_1 in ONE launch: both products accumulate in the same TMEM tile
                w0, w1 = weights
                ops.gemm_tf32(dys[0], w0[:, off:off + n_in], B, n_in, w0.shape[0], b_mn=True, out=dx, split_k=1,
                              a2=dys[1], b2=w1[:, off:off + n_in], K2=w1.shape[0])
            else:
                for i in range(n):  # dx[B, n_in] (+)= dy[B, H] . W[H, n_in]
                    w = weights[i]
                    ops.gemm_tf32(dys[i], w[:, off:off + n_in], B, n_in, w.shape[0], b_mn=True, out=dx, accumulate=i > 0,
                                  split_k=1)
        # piecewise optimizer sweep: the large weight blocks of this node are complete -> update them now, overlapping the
        # rest of backward (the sweep is HBM-bound, backward is not); it must follow every reader of the OLD weights, i.e.
        # the input-gradient product issued above
        bulk = None
        if early:
            # the sweeps go to the low-priority bulk stream, each behind its own gradient product and the readers of the
            # old weights
            bulk = ops.side_stream(x.device, BULK_STREAM)
            readers_done = torch.cuda.Event()
            readers_done.record(main)
            bulk.wait_event(readers_done)
            with torch.cuda.stream(bulk):
                for i in early:
                    bulk.wait_event(dw_done[i])
                    FUSED_OPT.early_block_update(weights_p[i])
        grads = []
        for i in range(n):
            if feats[i] is not None and ctx.lazy[i] is None and ctx.needs_input_grad[4 + 3 * i]:
                raise RuntimeError("multi_linear_big: the hand-made feature block is data (vae/encoder.py:250-287); it has no gradient")
            grads += [None, targets[i][0][1] if targets[i][0] is not None else None,
                      targets[i][1][1] if targets[i][1] is not None else None]
        if fork:
            # join now only if autograd receives a tensor produced on the side stream; otherwise at the end of backward
            if all(g is None for g in grads[1::3]) and all(g is None for g in grads[2::3]):
                for t in (x, *dys, *[f for f in feats if f is not None]):
                    t.record_stream(side)  # autograd may release them when this node returns; the fork still reads them
                defer_join(side)
            else:
                main.wait_stream(side)
        if bulk is not None:
            defer_join(bulk)
        return (dx, None, None, None, *grads)


def multi_linear_big(x: torch.Tensor, layers, n_in: int, col_offset: int = 0, bias_grad_elsewhere: bool = False,
                     streams=None):
    """``layers`` = [(nn.Linear, feat or LazyFeat or None), ...] applied to the same wide input x; returns a tuple of
    outputs.  x may carry padding columns beyond n_in (16-byte-aligned rows from the gather / BN kernels).
    ``bias_grad_elsewhere``: the layer feeds ``bn_act`` which also produces the Linear bias gradient (the column sums
    of its input gradient), so no separate reduction over dy is launched here.
    ``streams``: one CUDA stream (or None) per layer on which its forward product is issued."""
    if not _tma_addressable(x) and x.dim() == 2 and x.is_cuda and x.shape[0] * n_in <= (1 << 22):
        # a SMALL activation matrix whose row pitch is not a multiple of 16 bytes (z with --z-dim 6, say): zero-pad the rows
        # (differentiable); the wide [B, G] matrices always come from kernels that write padded rows
        x = torch.nn.functional.pad(x[:, :n_in], (0, (-n_in) % 4))
    _require_tma(x, "the input activations")
    args = []
    for lin, feat in layers:
        _require_tma(lin.weight, "an nn.Linear weight", col_offset)
        bias = lin.bias
        if bias is not None and bias_grad_elsewhere and torch.is_grad_enabled():
            bias = bias.detach()
        args += [feat, lin.weight, bias]
    return _MultiLinearBig.apply(x, n_in, col_offset, streams, *args)


def linear_big(x: torch.Tensor, lin: torch.nn.Linear, n_in: int, col_offset: int = 0, feat: Optional[torch.Tensor] = None,
               bias_grad_elsewhere: bool = False):
    return multi_linear_big(x, [(lin, feat)], n_in, col_offset, bias_grad_elsewhere)[0]


class _BN0Dropout(torch.autograd.Function):
    """dropout50(batchnorm0(x)) of EncodeNonZLatents (vae/encoder.py:239,279) through cb_bn0_forward / _backward."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, n_tracked, keep, training: bool, momentum: float, eps: float,
                n_genes: int):
        B, G = x.shape[0], n_genes
        y = ops.alloc_rows(B, G, x.device, zero=False)
        mean = torch.empty(G, device=x.device)
        rstd = torch.empty(G, device=x.device)
        lib().call("cb_bn0_forward", ptr(x), x.stride(0), B, G, ptr(gamma), ptr(beta), ptr(running_mean), ptr(running_var),
                   ptr(n_tracked) if training else None, ptr(keep), int(training), float(momentum), float(eps), ptr(y),
                   y.stride(0), ptr(mean), ptr(rstd), current_stream())
        ctx.save_for_backward(x, keep if keep is not None else torch.empty(0, device=x.device), mean, rstd)
        ctx.meta = (G, keep is not None)
        return y[:, :G]

    @staticmethod
    def backward(ctx, dy):
        x, keep, mean, rstd = ctx.saved_tensors
        G, has_keep = ctx.meta
        if dy.stride(1) != 1:
            dy = dy.contiguous()
        dgamma = torch.empty(G, device=x.device)
        dbeta = torch.empty(G, device=x.device)
        lib().call("cb_bn0_backward", ptr(x), x.stride(0), ptr(dy), dy.stride(0), x.shape[0], G, ptr(keep) if has_keep else None,
                   ptr(mean), ptr(rstd), ptr(dgamma), ptr(dbeta), current_stream())
        return None, dgamma, dbeta, None, None, None, None, None, None, None, None


def bn0_dropout(x: torch.Tensor, bn: torch.nn.BatchNorm1d, keep: Optional[torch.Tensor], training: bool, n_genes: int):
    """dropout(bn(x[:, :n_genes])) as a [B, n_genes] view of a 16-byte-row buffer (GEMM/TMA friendly)."""
    if not (x.is_cuda and x.dtype == torch.float32 and x.stride(1) == 1):
        raise RuntimeError("cellbender_b200 ops require fp32 CUDA tensors with unit column stride (no CPU fallback)")
    return _BN0Dropout.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked,
                             keep if training else None, training, bn.momentum if bn.momentum is not None else 0.1, bn.eps, n_genes)


ACT_CODE = {None: 0, "identity": 0, "softplus": 1, "relu": 2}


class _BNAct(torch.autograd.Function):
    """act(batchnorm(x)) for the [B, 512] hidden layers through cb_bn_act_forward / _backward (csrc/bn_act.cu).
    gamma / beta gradients -- and, when ``bias`` is given, the gradient of the Linear bias that produced x -- are
    written straight into the optimizer's flat gradient buffer."""

    @staticmethod
    def forward(ctx, x, gamma, beta, bn, act: int, training: bool, bias, feat, feat_weight):
        B, N = x.shape
        if x.stride(1) != 1:
            if feat is not None:
                raise RuntimeError("bn_act: the feature term is written back into x, which must have unit column stride")
            x = x.contiguous()
        y = torch.empty((B, N), dtype=torch.float32, device=x.device)
        mean = torch.empty(N, dtype=torch.float32, device=x.device)
        rstd = torch.empty(N, dtype=torch.float32, device=x.device)
        momentum = bn.momentum if bn.momentum is not None else 0.1
        n_feat = 0 if feat is None else feat.shape[1]
        lib().call("cb_bn_act_forward", ptr(x), x.stride(0), B, N, ptr(gamma), ptr(beta), ptr(bn.running_mean),
                   ptr(bn.running_var), ptr(bn.num_batches_tracked) if training else None, int(training), float(momentum),
                   float(bn.eps), act, ptr(y), y.stride(0), ptr(mean), ptr(rstd), ptr(feat), n_feat,
                   ptr(feat_weight) if feat is not None else None, feat_weight.stride(0) if feat is not None else 0,
                   current_stream())
        ctx.save_for_backward(x, mean, rstd)
        ctx.params = (gamma, beta, bias)
        ctx.meta = (act, training)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd = ctx.saved_tensors
        gamma, beta, bias = ctx.params
        act, training = ctx.meta
        B, N = x.shape
        if dy.stride(1) != 1:
            dy = dy.contiguous()
        dx = torch.empty((B, N), dtype=torch.float32, device=x.device)
        dg_buf, dg = _grad_target(gamma)
        db_buf, db = _grad_target(beta)
        dbias_buf = None
        if bias is not None and bias.requires_grad:
            dbias_buf, dbias = _grad_target(bias)
            if dbias is not None:  # no flat gradient view: hand the gradient to the parameter directly
                bias.grad = dbias
        lib().call("cb_bn_act_backward", ptr(x), x.stride(0), ptr(dy), dy.stride(0), B, N, ptr(gamma), ptr(beta), ptr(mean),
                   ptr(rstd), int(training), act, ptr(dx), dx.stride(0), ptr(dg_buf), ptr(db_buf), ptr(dbias_buf),
                   current_stream())
        return dx, dg, db, None, None, None, None, None, None


def bn_act(x: torch.Tensor, bn: torch.nn.BatchNorm1d, act: Optional[str], training: bool,
           bias_of: Optional[torch.nn.Linear] = None, feat: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``act(bn(x))`` in one kernel each way.  ``bias_of``: the nn.Linear that produced ``x`` when its bias gradient
    should come out of this op's backward (pair with ``bias_grad_elsewhere=True`` on that layer).
    ``feat`` [B, F]: x was produced WITHOUT the first F (feature) input columns of ``bias_of`` (multi_linear_big with a
    LazyFeat); their contribution feat @ bias_of.weight[:, :F]^T is added here, in place, before the normalisation."""
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2):
        raise RuntimeError("cellbender_b200 ops require 2-D fp32 CUDA tensors (no CPU fallback)")
    bias = bias_of.bias if (bias_of is not None and bias_of.bias is not None and torch.is_grad_enabled()) else None
    fw = None
    if feat is not None:
        assert bias_of is not None and feat.is_contiguous() and feat.shape[1] <= 8
        fw = bias_of.weight.detach()
    return _BNAct.apply(x, bn.weight, bn.bias, bn, ACT_CODE[act], training, bias, feat, fw)
