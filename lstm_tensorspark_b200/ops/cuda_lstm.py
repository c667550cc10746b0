"""CUDA LSTM layer ops = autograd.Functions around the hand-written kernels.

Fast path (bf16, H % 64 == 0, grid <= #SMs): hoisted input projection on the tcgen05 GEMM (csrc/gemm2_tcgen05.cu) + ONE
persistent tcgen05 kernel for the whole recurrence in each direction (csrc/lstm_seq_tcgen05.cu; weights resident in SMEM
up to H = 1024, streamed through the ring above).  Two stacked layers run as ONE layer wavefront (``_LSTMPairFn``: both
recurrences co-resident, the upper layer's x-projection / dX as a dataflow-gated GEMM on the idle SMs).  Generic path (any
shape / fp32): our CUDA-core GEMM per step (csrc/gemm_generic.cu) + the fused pointwise cell kernels (csrc/lstm_pointwise.cu).
Weight gradients are tcgen05 GEMMs over all T at once (``[4H, T·B] x [T·B, D | H]``, both operands MN-major and read in
place), fp32, written straight into the flat gradient buffer; bias gradients are deterministic column sums running next to
them.  Nothing in here reaches cuBLAS / cuDNN.  Math parity: /root/reference/src/models/recurrent/lstm.py:88-122.
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from .cuda_ext import ext
from . import cuda_gemm as G

_SYNC_WS = {}
_SM_COUNT = {}
FORCE_GENERIC = os.environ.get("LSTM_TS_FORCE_GENERIC", "0") == "1"
# Tuning / experiment knob of the persistent kernels (0 = defaults), bit fields as decoded in csrc/lstm_seq_tcgen05.cu seq_common():
#   [0:4) batch tiles per CTA (2 = one CTA alternates two tiles), [4:8) ring stages, bit 8 force streamed weights,
#   [12:15) timing-only debug mode (1 skip loads, 2 skip MMAs, 3 in-order stream, 4 half-size loads, 5 no bookkeeping stores,
#   6 no L2 prefetch, 7 cluster-scope acquire on the exchange barriers), [16:18) sync mode (0 per-k-block dataflow counters,
#   1 one counter per batch tile = grid barrier, 2 per-CTA flags), bit 18 acquire polls, bit 19 no forward K-split.
SEQ_VARIANT = int(os.environ.get("LSTM_TS_SEQ_VARIANT", "0"))
STATS = {"fast_fwd": 0, "fast_bwd": 0, "generic_fwd": 0, "generic_bwd": 0, "tc_gemm": 0, "kernels": 0}


# Gradient-bucket overlap (engine.TrainEngine + parallel/fused_comm.py): HOOKS["grads_written"] is called after every weight /
# bias gradient has been written; callables queued in AFTER_SEQ_BWD are run right after the NEXT big backward kernel has been
# launched (a persistent recurrence kernel or a weight-gradient GEMM: both execute griddepcontrol.launch_dependents).  They
# launch a finished bucket's fused allreduce + update as a programmatic dependent, so it runs NEXT TO that kernel (on the SMs
# the recurrence leaves idle / co-resident with the GEMM's CTAs) instead of after it.
HOOKS = {"grads_written": None}
AFTER_SEQ_BWD = []          # [(generation when queued, closure)]
_GEN = {"n": 0}


def queue_after_big_launch(fn):
    AFTER_SEQ_BWD.append((_GEN["n"], fn))


def _big_launch_begin():
    """A big backward kernel (recurrence / weight-gradient GEMM) is about to be launched as an ORDINARY launch: everything
    enqueued before it is complete when it starts."""
    _GEN["n"] += 1


def _after_big_launch(flush: bool = False):
    """Launch the queued bucket closures as programmatic dependents of the kernel just launched - but only those queued
    BEFORE that kernel was launched: a dependent may start while its primary runs, so its inputs must not come from it."""
    while AFTER_SEQ_BWD and (flush or AFTER_SEQ_BWD[0][0] < _GEN["n"]):
        AFTER_SEQ_BWD.pop(0)[1]()


def _grads_written():
    h = HOOKS["grads_written"]
    if h is not None and not _CHUNKING["on"]:
        h()

_PARAMS = {}          # fp32 param address -> (bf16 shadow view, fp32 grad view), maintained by models.flat.FlatParams
DIRECT_GRADS = os.environ.get("LSTM_TS_DIRECT_GRADS", "1") == "1"


def register_param(addr: int, shadow: torch.Tensor, grad: torch.Tensor, owner=None) -> None:
    import weakref
    _PARAMS[addr] = (shadow, grad, weakref.ref(owner) if owner is not None else None)


def _lookup(addr: int):
    ent = _PARAMS.get(addr)
    if ent is None:
        return None
    if ent[2] is not None and ent[2]() is None:       # the FlatParams buffer died: its address may have been reused
        del _PARAMS[addr]
        return None
    return ent


def _lowp(w: torch.Tensor, cd: torch.dtype) -> torch.Tensor:
    """bf16 copy of a weight: the optimizer-maintained shadow when there is one, a cast otherwise."""
    if cd == torch.bfloat16:
        ent = _lookup(w.data_ptr())
        if ent is not None and ent[0].shape == w.shape:
            return ent[0]
    return w.detach().to(cd).contiguous()


def grad_sink(w_addr: int):
    """-> (fp32 grad view inside the flat buffer, accumulate flag) for a registered parameter, or None.
    accumulate False = first write of this step: the kernel overwrites (no zero-filled buffer needed)."""
    ent = _lookup(w_addr) if DIRECT_GRADS else None
    if ent is None:
        return None
    owner = ent[2]() if ent[2] is not None else None
    if owner is None or w_addr not in owner._direct:
        if owner is not None:
            owner.ensure_zeroed(w_addr)
        return ent[1], True
    return ent[1], owner.take_sink(w_addr)


def _accumulate_grad(w_addr: int, a_t: torch.Tensor, b: torch.Tensor, b_folded: bool = False):
    """dW = a_t @ b in fp32 (``a_t`` = dG^T as a transposed view, ``b`` = the layer input: both operands MN-major, read in
    place by the tcgen05 GEMM; ``b_folded``: ``b`` is the batch-major [B,T,D] array standing for the time-major [T*B, D]
    matrix).  When the parameter lives in a FlatParams buffer the product lands straight in its grad
    view (overwrite on the first write of a step, accumulate afterwards) and None is returned to autograd."""
    ops = dict(a=a_t, b_t=None, b_folded=b) if b_folded else dict(a=a_t, b_t=b.t())
    sink = grad_sink(w_addr)
    if sink is not None:
        _big_launch_begin()
        G.matmul(out=sink[0], accumulate=sink[1], **ops)
        _after_big_launch()                  # finished buckets of earlier gradients: allreduce them under this GEMM
        _grads_written()
        return None
    return G.matmul(out_dtype=torch.float32, **ops)


_BIAS_SPLIT = {}


def _bias_grad(b_addr: int, dg2d: torch.Tensor, under_gemm: bool = False, part: int = -1):
    """db = column sums of dG; straight into the flat grad view when there is one.  ``under_gemm``: the previous launch of the
    stream is a weight-gradient GEMM over the same dG that leaves SMs idle - run next to it (programmatic dependent launch).
    ``part`` 0 / 1: only the first / second half of the columns (one half under each of the layer's two weight-gradient GEMMs:
    on the ~20 idle SMs a half takes about as long as the GEMM it hides under); the value for autograd comes from part 1."""
    fast = dg2d.is_cuda and dg2d.dtype == torch.bfloat16 and dg2d.shape[1] % 512 == 0 and dg2d.is_contiguous()
    if part == 0:
        if not fast:
            return None                                   # everything happens with the part-1 call
        sink = grad_sink(b_addr)
        _BIAS_SPLIT[b_addr] = sink
        if sink is None:
            return None
        half = dg2d.shape[1] // 2
        STATS["kernels"] += 1
        ext().colsum_bf16_into(dg2d, sink[0], not sink[1], under_gemm, 0, half)
        return None
    if part == 1 and fast and b_addr in _BIAS_SPLIT:
        sink = _BIAS_SPLIT.pop(b_addr)
        if sink is not None:
            half = dg2d.shape[1] // 2
            STATS["kernels"] += 1
            ext().colsum_bf16_into(dg2d, sink[0], not sink[1], under_gemm, half, half)
            _grads_written()
            return None
        return ext().colsum_bf16(dg2d)
    fast = dg2d.is_cuda and dg2d.dtype == torch.bfloat16 and dg2d.shape[1] % 256 == 0 and dg2d.is_contiguous()
    sink = grad_sink(b_addr)
    if fast:
        STATS["kernels"] += 1
        if sink is not None:
            ext().colsum_bf16_into(dg2d, sink[0], not sink[1], under_gemm and part < 0)
            _grads_written()
            return None
        return ext().colsum_bf16(dg2d)
    ones = torch.ones(1, dg2d.shape[0], dtype=dg2d.dtype, device=dg2d.device)
    if sink is not None:
        G.matmul(ones, dg2d.t(), out=sink[0].view(1, -1), accumulate=sink[1])
        _grads_written()
        return None
    return G.matmul(ones, dg2d.t(), out_dtype=torch.float32).view(-1)


SYNC_WORDS = 8192        # csrc/lstm_seq_tcgen05.cu kSyncWords; the last word is the sticky error flag


def _sync_ws(device) -> torch.Tensor:
    key = (device.type, device.index)
    if key not in _SYNC_WS:
        _SYNC_WS[key] = torch.zeros(SYNC_WORDS, dtype=torch.int32, device=device)
    return _SYNC_WS[key]


def check_kernel_errors(device) -> None:
    """Raise if a persistent kernel hit its bounded-spin timeout (sticky flag, costs one D2H read)."""
    ws = _SYNC_WS.get((device.type, device.index))
    if ws is not None and int(ws[SYNC_WORDS - 1].item()) != 0:
        raise RuntimeError("lstm_seq kernel aborted: an in-kernel wait timed out (see csrc/lstm_seq_tcgen05.cu)")
    for (di, _tag), ent in list(globals().get("_WS_PAIR", {}).items()):
        if di == device.index and (int(ent[SYNC_WORDS - 1].item()) != 0 or int(ent[2 * SYNC_WORDS - 1].item()) != 0):
            raise RuntimeError("lstm_seq kernel (layer wavefront) aborted: an in-kernel wait timed out")


def _sms(device) -> int:
    key = device.index
    if key not in _SM_COUNT:
        _SM_COUNT[key] = torch.cuda.get_device_properties(device).multi_processor_count
    return _SM_COUNT[key]


_CORES = {}


def _coresident_ctas(device) -> int:
    """CTAs of the persistent kernels that can be co-resident: the backward kernel runs in clusters of 4 (measured on
    B200: 33 clusters = 132 CTAs of 148 SMs), the forward K-split in clusters of 2."""
    key = device.index
    if key not in _CORES:
        n = _sms(device)
        try:
            with torch.cuda.device(device):
                c4 = int(ext().lstm_seq_cluster_probe(4))
            if c4 > 0:
                n = min(n, 4 * c4)
        except Exception:                                   # noqa: BLE001
            pass
        _CORES[key] = n
    return _CORES[key]


def fast_path_supported(B: int, H: int, dtype: torch.dtype, device) -> bool:
    if FORCE_GENERIC or dtype != torch.bfloat16 or H % 64 != 0:
        return False
    tiles_m = (B + 127) // 128
    # one CTA per (batch tile, 64 gate columns); all of them must be co-resident (dataflow sync between CTAs), in clusters
    # of 4 for the backward pass.  The weight slice stays in shared memory when it fits (H <= 1024), larger H streams it
    # through the ring (csrc/lstm_seq_tcgen05.cu, kStream)
    return tiles_m * (H // 16) <= _coresident_ctas(device) and tiles_m <= 16


def _batch_chunk(B: int, H: int, dtype: torch.dtype, device) -> Optional[int]:
    """Largest multiple of 128 rows whose CTAs fit the device, when the whole batch does not (else None)."""
    if FORCE_GENERIC or dtype != torch.bfloat16 or H % 64 != 0 or fast_path_supported(B, H, dtype, device):
        return None
    tiles = _coresident_ctas(device) // (H // 16)
    if tiles < 1:
        return None
    chunk = min(tiles, 16) * 128
    return chunk if chunk < B else None


def _mm_f32(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a [M,K] @ b [K,N] -> fp32 (own kernels: tcgen05 when bf16 and aligned, CUDA-core GEMM otherwise)."""
    return G.matmul(a, b.t(), out_dtype=torch.float32)


def _transposed(w: torch.Tensor) -> torch.Tensor:
    """``w [R,C]`` -> contiguous ``[C,R]`` (tile-transpose kernel for 16-bit CUDA tensors)."""
    if w.is_cuda and w.dim() == 2 and w.element_size() == 2 and w.is_contiguous():
        STATS["kernels"] += 1
        return ext().transpose2d(w)
    return w.t().contiguous()


def _gemm_tn(a: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """a [M,K] @ w[N,K]^T -> [M,N] in a.dtype."""
    STATS["tc_gemm"] += 1
    STATS["kernels"] += 1
    return G.matmul(a, w, out_dtype=a.dtype)


_CHUNKING = {"on": False}      # True while lstm_layer_sequence feeds batch chunks (weight gradients then accumulate over calls)


class _LSTMSeqFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_seq, h0, c0, w_x, w_h, bias):
        E = ext()
        T, B, D = x_seq.shape
        H = w_h.shape[1]
        cd = x_seq.dtype
        x2d = x_seq.reshape(T * B, D).contiguous()
        w_x_c = _lowp(w_x, cd)
        w_h_c = _lowp(w_h, cd)
        bias_f = bias.detach().float().contiguous()
        gx = _gemm_tn(x2d, w_x_c).view(T, B, 4 * H)
        fast = fast_path_supported(B, H, cd, x_seq.device)
        c0f = c0.detach().float().contiguous()
        h0c = h0.detach().to(cd).contiguous()
        if fast:
            h_seq, c_seq, act = E.lstm_seq_fwd(gx, w_h_c, bias_f, h0c, c0f, _sync_ws(x_seq.device), SEQ_VARIANT)
            STATS["fast_fwd"] += 1
            STATS["kernels"] += 1
        else:
            h_seq = torch.empty(T + 1, B, H, dtype=cd, device=x_seq.device)
            c_seq = torch.empty(T + 1, B, H, dtype=torch.float32, device=x_seq.device)
            act = torch.empty(T, B, 4 * H, dtype=cd, device=x_seq.device)
            h_seq[0].copy_(h0c)
            c_seq[0].copy_(c0f)
            pre = torch.empty(B, 4 * H, dtype=cd, device=x_seq.device)
            for t in range(T):
                pre.copy_(gx[t])
                G.matmul(h_seq[t], w_h_c, out=pre, accumulate=True)       # pre = gx[t] + h_{t-1} W_h^T
                h, c, a = E.lstm_pointwise_fwd(pre, bias_f, c_seq[t])
                h_seq[t + 1].copy_(h)
                c_seq[t + 1].copy_(c)
                act[t].copy_(a)
            STATS["generic_fwd"] += 1
            STATS["kernels"] += T
        ctx.save_for_backward(x2d, h_seq, c_seq, act, w_x_c, w_h_c)
        ctx.set_materialize_grads(False)       # an unused output must arrive as None, not as a zero-filled [T,B,H] tensor
        ctx.fast = fast
        ctx.whole_batch = not _CHUNKING["on"]
        ctx.dims = (T, B, D, H)
        ctx.w_addrs = (w_x.data_ptr(), w_h.data_ptr(), bias.data_ptr())
        ctx.in_dtypes = (h0.dtype, c0.dtype)
        # h_T is its own output (not a slice of the first one taken by the caller): a consumer of the final state only - the
        # classifier on top of the stack - then sends back a [B,H] gradient instead of a zero-filled [T,B,H] one
        return h_seq[1:], h_seq[T], c_seq[T]

    @staticmethod
    def backward(ctx, dh_seq, dh_T, dc_T):
        E = ext()
        x2d, h_seq, c_seq, act, w_x_c, w_h_c = ctx.saved_tensors
        T, B, D, H = ctx.dims
        cd = act.dtype
        dev = act.device
        if dh_seq is not None:
            dh_seq = dh_seq.to(cd).contiguous()
        elif not ctx.fast:
            dh_seq = torch.zeros(T, B, H, dtype=cd, device=dev)
        dcT = (dc_T.float().contiguous() if dc_T is not None else torch.zeros(B, H, dtype=torch.float32, device=dev))
        dhT = (dh_T.float().contiguous() if dh_T is not None else torch.zeros(B, H, dtype=torch.float32, device=dev))
        if ctx.fast:
            w_hT = _transposed(w_h_c)
            _big_launch_begin()
            dpre, dh0, dc0 = E.lstm_seq_bwd(dh_seq, w_hT, act, c_seq, dhT, dcT, _sync_ws(dev), SEQ_VARIANT)
            STATS["fast_bwd"] += 1
            STATS["kernels"] += 1
            _after_big_launch()                  # finished gradient buckets of the layers above: sync them under this recurrence
        else:
            dpre = torch.empty_like(act)
            dh_rec: Optional[torch.Tensor] = dhT if dh_T is not None else None
            dc = dcT
            for t in range(T - 1, -1, -1):
                dp, dc = E.lstm_pointwise_bwd(dh_seq[t], dh_rec, dc, act[t], c_seq[t], c_seq[t + 1])
                dpre[t].copy_(dp)
                dh_rec = _mm_f32(dp, w_h_c)
            dh0, dc0 = dh_rec, dc
            STATS["generic_bwd"] += 1
            STATS["kernels"] += T
        dg2d = dpre.view(T * B, 4 * H)
        dg_t = dg2d.t()
        dw_x = _accumulate_grad(ctx.w_addrs[0], dg_t, x2d)
        dw_h = _accumulate_grad(ctx.w_addrs[1], dg_t, h_seq[:T].reshape(T * B, H))
        db = _bias_grad(ctx.w_addrs[2], dg2d)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = G.matmul(dg2d, w_x_c.t(), out_dtype=cd).view(T, B, D)        # dG · W_x: W_x read in place as an MN-major operand
            STATS["kernels"] += 1
        h0_dt, c0_dt = ctx.in_dtypes
        return dx, dh0.to(h0_dt), dc0.to(c0_dt), dw_x, dw_h, db


def lstm_layer_sequence(x_seq, h0, c0, w_x, w_h, bias):
    """``x_seq [T,B,D]`` (bf16 or fp32) -> ``(h_seq [T,B,H], h_T, c_T)``."""
    if (not x_seq.is_contiguous() and not x_seq.requires_grad and x_seq.transpose(0, 1).is_contiguous()
            and (x_seq.shape[2] * x_seq.element_size()) % 16 == 0):
        x_seq = ext().transpose01(x_seq.transpose(0, 1))     # batch-major feed -> time-major, a row permutation at copy speed
        STATS["kernels"] += 1
    T, B, _ = x_seq.shape
    H = w_h.shape[1]
    chunk = _batch_chunk(B, H, x_seq.dtype, x_seq.device)
    if chunk is not None:
        # more batch tiles than the persistent kernels can keep co-resident: the sequences are independent, so run the fast
        # path per batch chunk (weight-gradient contributions accumulate across chunks) instead of the per-step generic path
        _CHUNKING["on"] = True
        try:
            outs = [_LSTMSeqFn.apply(x_seq[:, b0:b0 + chunk].contiguous(), h0[b0:b0 + chunk], c0[b0:b0 + chunk], w_x, w_h, bias)
                    for b0 in range(0, B, chunk)]
        finally:
            _CHUNKING["on"] = False
        STATS["batch_chunks"] = STATS.get("batch_chunks", 0) + len(outs)
        return (torch.cat([o[0] for o in outs], dim=1), torch.cat([o[1] for o in outs], dim=0),
                torch.cat([o[2] for o in outs], dim=0))
    return _LSTMSeqFn.apply(x_seq.contiguous(), h0, c0, w_x, w_h, bias)


# =====================================================================================================================
# Layer wavefront: two stacked layers' recurrences run CO-RESIDENT (64 + 64 CTAs, two batch tiles per CTA over the same resident
# weight slice), chained through a dataflow-gated tcgen05 GEMM on the ~20 SMs they leave idle:
#     forward :  L_a step t  ->  gx_b[t] = h_a[t] W_xb^T (gated GEMM)  ->  L_b step t
#     backward:  L_b step t  ->  dh_a[t] = dG_b[t] W_xb  (gated GEMM)  ->  L_a step t
# The reference stacks layers strictly one after the other (/root/reference/src/models/recurrent/rnn.py:38-42); here layer l+1
# trails layer l by a couple of time steps and the next layer's input projection leaves the critical path altogether.
# =====================================================================================================================
FOLDED_FEED = os.environ.get("LSTM_TS_FOLDED_FEED", "1") != "0"   # batch-major input read in place by the first layer's GEMMs
WAVEFRONT = os.environ.get("LSTM_TS_WAVEFRONT", "1") == "1"
_SIDE_STREAMS = {}
_WS_PAIR = {}


def _side_streams(device):
    key = device.index
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = (torch.cuda.Stream(device=device), torch.cuda.Stream(device=device))
    return _SIDE_STREAMS[key]


_WARM = set()


def _warm_wavefront_kernels(device):
    """CUDA loads kernels lazily, and loading one may wait for every running kernel to finish.  The wavefront's kernels WAIT
    FOR EACH OTHER on the device, so a kernel that is loaded for the first time while its producers are already spinning would
    deadlock (until the bounded spins time out).  Load the gated GEMM instantiations once, before the first concurrent use."""
    if device.index in _WARM:
        return
    a = torch.zeros(256, 64, dtype=torch.bfloat16, device=device)
    w = torch.zeros(256, 64, dtype=torch.bfloat16, device=device)
    wt = torch.zeros(64, 256, dtype=torch.bfloat16, device=device)
    ext().gemm2(a, w, ctas=1, bn=256)
    ext().gemm2(a, wt, b_mn=True, ctas=1, bn=256)
    torch.cuda.synchronize(device)
    _WARM.add(device.index)


def _pair_ws(device, tag: str, n_done: int):
    """[sync ws of the head kernel | sync ws of the tail kernel | completion counters of the gated GEMM] in one allocation."""
    key = (device.index, tag)
    ent = _WS_PAIR.get(key)
    if ent is None or ent.numel() < 2 * SYNC_WORDS + n_done:
        ent = torch.zeros(2 * SYNC_WORDS + max(n_done, 1), dtype=torch.int32, device=device)
        _WS_PAIR[key] = ent
    return ent[:SYNC_WORDS], ent[SYNC_WORDS:2 * SYNC_WORDS], ent[2 * SYNC_WORDS:2 * SYNC_WORDS + n_done]


def wavefront_supported(x_seq: torch.Tensor, h_a: int, h_b: int) -> bool:
    """Two co-resident layers need: bf16 fast path, B = 256 (two batch tiles per CTA, one GEMM tile per time step), resident
    weights (H <= 1024), 256-aligned widths, and 2 * H/16 CTAs + a few GEMM CTAs within the device."""
    if not (WAVEFRONT and x_seq.is_cuda and x_seq.dtype == torch.bfloat16 and not FORCE_GENERIC and x_seq.dim() == 3):
        return False
    T, B, D = x_seq.shape
    if B != 256 or T < 2 or D % 8 != 0 or (SEQ_VARIANT & 15) > 2:
        return False
    for h in (h_a, h_b):
        if h % 256 != 0 or h > 1024:
            return False
    return h_a // 16 + h_b // 16 + 8 <= _coresident_ctas(x_seq.device) + 16 and h_a // 16 + h_b // 16 + 8 <= _sms(x_seq.device)


WAVE_SYNC_MODE = int(os.environ.get("LSTM_TS_WAVE_SYNC", "1"))      # 1: one arrival counter per batch tile (measured 2.5 % faster with two
                                                                    # tiles per CTA), 0: one per operand k-block (profiles/logs/tiles2_tune.log)


def _wave_variant() -> int:
    return (SEQ_VARIANT & ~(15 | (3 << 16))) | 2 | ((WAVE_SYNC_MODE & 1) << 16)


def _gate_off(var: int) -> int:
    return 0 if (var >> 16) & 3 == 1 else 512


def _gate_cfg(var: int, tiles_m: int, nkb: int, per_kb_step: int, ctas_per_tile: int, base_steps: int, step_sign: int, rows: int, bwd: bool):
    """Gate of the wavefront GEMM on the producer layer's arrival counters: [count, stride, base, per_step, rows_per_step, use_last,
    reverse_m].  A time step's natural-layout rows are complete with the producer's (step + 1)-th signal, hence base = 2 steps
    (forward: target(t) = per_step * (t + 2)) or T + 1 (backward: target(t) = per_step * (T + 1 - t))."""
    if (var >> 16) & 3 == 1:                # one counter per batch tile, every CTA of the tile arrives once per step
        per = ctas_per_tile
        return [tiles_m, 1, per * base_steps, per * step_sign, rows, 0 if bwd else 1, 1 if bwd else 0]
    per = per_kb_step                       # one counter per operand k-block (fwd: 4 producer CTAs, bwd: 1)
    return [tiles_m * nkb, 32, per * base_steps, per * step_sign, rows, 0 if bwd else 1, 1 if bwd else 0]


class _LSTMPairFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_seq, h0a, c0a, w_xa, w_ha, b_a, h0b, c0b, w_xb, w_hb, b_b):
        E = ext()
        dev = x_seq.device
        T, B, D = x_seq.shape
        Ha, Hb = w_ha.shape[1], w_hb.shape[1]
        cd = x_seq.dtype
        # a batch-major input ([B,T,D] storage behind a transposed view) is read in place by the x-projection and by the
        # weight-gradient GEMM of the first layer (folded tensor map, csrc/gemm2_tcgen05.cu): no transpose pass
        x_bm = x_seq.transpose(0, 1) if not x_seq.is_contiguous() else None
        x2d = x_seq.reshape(T * B, D) if x_bm is None else x_bm
        wxa, wha, wxb, whb = _lowp(w_xa, cd), _lowp(w_ha, cd), _lowp(w_xb, cd), _lowp(w_hb, cd)
        ba_f, bb_f = b_a.detach().float().contiguous(), b_b.detach().float().contiguous()
        h0a_c, h0b_c = h0a.detach().to(cd).contiguous(), h0b.detach().to(cd).contiguous()
        c0a_f, c0b_f = c0a.detach().float().contiguous(), c0b.detach().float().contiguous()
        _warm_wavefront_kernels(dev)
        if x_bm is None:
            gx_a = _gemm_tn(x2d, wxa).view(T, B, 4 * Ha)
        else:
            STATS["tc_gemm"] += 1; STATS["kernels"] += 1; STATS["folded_feed"] = STATS.get("folded_feed", 0) + 1
            gx_a = G.matmul(None, wxa, out_dtype=cd, a_folded=x_bm).view(T, B, 4 * Ha)
        opt = dict(dtype=cd, device=dev)
        h_seq_a = torch.empty(T + 1, B, Ha, **opt); c_seq_a = torch.empty(T + 1, B, Ha, dtype=torch.float32, device=dev)
        act_a = torch.empty(T, B, 4 * Ha, **opt); til_a = torch.empty((T + 1) * 2 * 128 * Ha, **opt)
        h_seq_b = torch.empty(T + 1, B, Hb, **opt); c_seq_b = torch.empty(T + 1, B, Hb, dtype=torch.float32, device=dev)
        act_b = torch.empty(T, B, 4 * Hb, **opt); til_b = torch.empty((T + 1) * 2 * 128 * Hb, **opt)
        gx_b = torch.empty(T, B, 4 * Hb, **opt)
        tn = 4 * Hb // 256
        ws_a, ws_b, done = _pair_ws(dev, "fwd", T * tn * 2)
        done.zero_()                                                     # (the prologue kernels zero ws_a / ws_b)
        var = _wave_variant()                                            # two batch tiles per CTA: 64 CTAs per layer at H = 1024
        # ONE stream, a programmatic-dependent-launch chain: L_a -> L_b (starts once every CTA of L_a is resident) -> gated GEMM
        # (starts once every CTA of L_b is resident, on the SMs that are left).  The order in which the three grids take their
        # SMs is thereby fixed (a kernel that is still queueing could otherwise starve the chain head of co-resident SMs).
        # Everything the later kernels need up front (prologues, zeroed counters) is enqueued before the chain head.
        E.lstm_seq_prologue(h0a_c, c0a_f, h_seq_a, c_seq_a, til_a, ws_a)
        E.lstm_seq_prologue(h0b_c, c0b_f, h_seq_b, c_seq_b, til_b, ws_b)
        E.lstm_seq_fwd_into(gx_a, wha, ba_f, h0a_c, c0a_f, h_seq_a, c_seq_a, act_a, til_a, ws_a, var, None, 0, True, 0, 1)
        E.lstm_seq_fwd_into(gx_b, whb, bb_f, h0b_c, c0b_f, h_seq_b, c_seq_b, act_b, til_b, ws_b, var, done, tn, False, 0, 3)
        # single-CTA tiles: the recurrences' CTAs are spread one per TPC, the SMs they leave free rarely form CTA pairs
        free_ctas = max(1, _sms(dev) - Ha // 16 - Hb // 16)
        E.gemm2(h_seq_a[1:].view(T * B, Ha), wxb, out=gx_b.view(T * B, 4 * Hb), ctas=1, bn=256, max_ctas=free_ctas,
                gate=ws_a[_gate_off(var):], gate_cfg=_gate_cfg(var, 2, Ha // 64, 4, 4 * Ha // 64, 2, 1, B, False), done=done,
                gate_err=ws_a[SYNC_WORDS - 1:], pdl=True)
        STATS["fast_fwd"] += 2
        STATS["kernels"] += 3
        STATS["wavefront_fwd"] = STATS.get("wavefront_fwd", 0) + 1
        ctx.save_for_backward(x2d, h_seq_a, c_seq_a, act_a, h_seq_b, c_seq_b, act_b, wxa, wha, wxb, whb)
        ctx.set_materialize_grads(False)
        ctx.dims = (T, B, D, Ha, Hb)
        ctx.x_folded = x_bm is not None
        ctx.addrs = (w_xa.data_ptr(), w_ha.data_ptr(), b_a.data_ptr(), w_xb.data_ptr(), w_hb.data_ptr(), b_b.data_ptr())
        ctx.in_dtypes = (h0a.dtype, c0a.dtype, h0b.dtype, c0b.dtype)
        return h_seq_b[1:], h_seq_a[T], c_seq_a[T], h_seq_b[T], c_seq_b[T]

    @staticmethod
    def backward(ctx, dh_seq_b, dhT_a, dcT_a, dhT_b, dcT_b):
        E = ext()
        x2d, h_seq_a, c_seq_a, act_a, h_seq_b, c_seq_b, act_b, wxa, wha, wxb, whb = ctx.saved_tensors
        T, B, D, Ha, Hb = ctx.dims
        cd, dev = act_a.dtype, act_a.device
        if dh_seq_b is not None:
            dh_seq_b = dh_seq_b.to(cd).contiguous()
        f32 = dict(dtype=torch.float32, device=dev)
        z = lambda g, H: g.float().contiguous().clone() if g is not None else torch.zeros(B, H, **f32)
        dh0a, dc0a, dh0b, dc0b = z(dhT_a, Ha), z(dcT_a, Ha), z(dhT_b, Hb), z(dcT_b, Hb)
        whT_a, whT_b = _transposed(wha), _transposed(whb)
        dpre_a = torch.empty_like(act_a); dpre_b = torch.empty_like(act_b)
        til_a = torch.empty(T * 2 * 128 * 4 * Ha, dtype=cd, device=dev); til_b = torch.empty(T * 2 * 128 * 4 * Hb, dtype=cd, device=dev)
        dx_b = torch.empty(T, B, Ha, dtype=cd, device=dev)                # = the gradient into every h_a[t]
        tn = Ha // 256
        ws_b, ws_a, done = _pair_ws(dev, "bwd", T * tn * 2)               # head of the backward chain is layer b
        ws_a[:SYNC_WORDS - 1].zero_(); ws_b[:SYNC_WORDS - 1].zero_(); done.zero_()
        var = _wave_variant()
        # programmatic-dependent-launch chain on one stream (see forward): L_b -> L_a -> gated dX GEMM
        E.lstm_seq_bwd_into(dh_seq_b, whT_b, act_b, c_seq_b, dpre_b, dh0b, dc0b, til_b, ws_b, var, None, 0, True, 0, 1)
        E.lstm_seq_bwd_into(dx_b, whT_a, act_a, c_seq_a, dpre_a, dh0a, dc0a, til_a, ws_a, var, done, tn, False, 0, 3)
        free_ctas = max(1, _sms(dev) - Ha // 16 - Hb // 16)
        E.gemm2(dpre_b.view(T * B, 4 * Hb), wxb, out=dx_b.view(T * B, Ha), b_mn=True, ctas=1, bn=256, max_ctas=free_ctas,
                gate=ws_b[_gate_off(var):], gate_cfg=_gate_cfg(var, 2, 4 * Hb // 64, 1, 4 * Hb // 64, T + 1, -1, B, True), done=done,
                gate_err=ws_b[SYNC_WORDS - 1:], pdl=True)
        STATS["fast_bwd"] += 2
        STATS["kernels"] += 5
        a = ctx.addrs
        dg_b = dpre_b.view(T * B, 4 * Hb)
        # the bias column sums run NEXT TO the first weight-gradient GEMM of their layer (same dG, idle SMs), not after it
        # (each GEMM is followed by: finished gradient buckets [programmatic dependents of the GEMM], then half of the layer's
        # bias column sums [programmatic dependent of whatever was launched last] - all three run side by side)
        dw_xb = _accumulate_grad(a[3], dg_b.t(), h_seq_a[1:].reshape(T * B, Ha))
        _bias_grad(a[5], dg_b, under_gemm=dw_xb is None, part=0)
        dw_hb = _accumulate_grad(a[4], dg_b.t(), h_seq_b[:T].reshape(T * B, Hb))
        db_b = _bias_grad(a[5], dg_b, under_gemm=dw_hb is None, part=1)
        dg_a = dpre_a.view(T * B, 4 * Ha)
        dw_xa = _accumulate_grad(a[0], dg_a.t(), x2d, b_folded=ctx.x_folded)
        _bias_grad(a[2], dg_a, under_gemm=dw_xa is None, part=0)
        dw_ha = _accumulate_grad(a[1], dg_a.t(), h_seq_a[:T].reshape(T * B, Ha))
        db_a = _bias_grad(a[2], dg_a, under_gemm=dw_ha is None, part=1)
        dx = None
        if ctx.needs_input_grad[0]:                                       # (never with a folded input: lstm_pair_sequence)
            dx = G.matmul(dg_a, wxa.t(), out_dtype=cd).view(T, B, D)
            STATS["kernels"] += 1
        t = ctx.in_dtypes
        return dx, dh0a.to(t[0]), dc0a.to(t[1]), dw_xa, dw_ha, db_a, dh0b.to(t[2]), dc0b.to(t[3]), dw_xb, dw_hb, db_b


def lstm_pair_sequence(x_seq, la, lb):
    """Two stacked layers as one wavefront op.  ``la`` / ``lb`` = (h0, c0, w_x, w_h, bias).  -> (h_seq_b, hT_a, cT_a, hT_b, cT_b)."""
    if (not x_seq.is_contiguous() and not x_seq.requires_grad and x_seq.transpose(0, 1).is_contiguous()
            and (x_seq.shape[2] * x_seq.element_size()) % 16 == 0):
        if FOLDED_FEED and G.folded_ok(x_seq.transpose(0, 1)):
            return _LSTMPairFn.apply(x_seq, *la, *lb)                     # read in place (see forward)
        x_seq = ext().transpose01(x_seq.transpose(0, 1))
        STATS["kernels"] += 1
    return _LSTMPairFn.apply(x_seq.contiguous(), *la, *lb)
