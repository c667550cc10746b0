"""Host-side logic of the GPU path that does not need a GPU: lazy gradient zeroing, ownership of Adam slots under bucketed
two-shot allreduce, the wavefront GEMM's gate configuration, the generation-gated bucket queue."""
import types

import torch

from lstm_tensorspark_b200.models.flat import FlatParams


def _flat():
    ps = [torch.nn.Parameter(torch.randn(8, 4)), torch.nn.Parameter(torch.randn(8, 8)), torch.nn.Parameter(torch.randn(8))]
    other = [torch.nn.Parameter(torch.randn(8, 3))]
    return FlatParams(ps, other), ps, other


def test_lazy_gradient_zeroing_protocol():
    flat, ps, other = _flat()
    flat.grad.fill_(7.0)
    flat.enable_direct_grads(ps)                       # w_x, w_h, bias are written by kernels; `other` goes through autograd
    flat.zero_grad()
    assert float(other[0].grad.abs().sum()) == 0.0     # autograd-accumulated parameters are really zeroed ...
    assert float(ps[0].grad.min()) == 7.0              # ... direct ones are only marked stale (no memset)
    a0 = ps[0].data_ptr()
    assert flat.take_sink(a0) is False                 # first write of the step: overwrite
    assert flat.take_sink(a0) is True                  # later writes: accumulate
    flat.ensure_zeroed(ps[1].data_ptr())               # a gradient about to be accumulated by autograd: zeroed on demand
    assert float(ps[1].grad.abs().sum()) == 0.0
    flat.finalize_grads()                              # parameters that got no gradient this step hold zeros afterwards
    assert float(ps[2].grad.abs().sum()) == 0.0 and not flat._stale
    assert float(ps[0].grad.min()) == 7.0              # (written by "the kernel": untouched)
    flat.zero_grad()
    assert flat.take_sink(a0) is False                 # next step starts over


def test_direct_set_follows_a_rebase():
    flat, ps, other = _flat()
    flat.enable_direct_grads(ps)
    old = set(flat._direct)
    flat.rebase(torch.zeros_like(flat.data), torch.zeros_like(flat.grad))
    assert flat._direct == {p.data_ptr() for p in ps} and flat._direct != old


def test_owned_ranges_of_bucketed_two_shot_adam():
    from lstm_tensorspark_b200.parallel.fused_comm import FusedComm
    n = 4 * 1000
    for world in (2, 3, 8):
        covered = torch.zeros(2 * n, dtype=torch.int32)
        for rank in range(world):
            fake = types.SimpleNamespace(rank=rank, world_size=world, _state_buckets=[(0, n, True), (n, 2 * n, False)])
            for lo, hi in FusedComm._owned_ranges(fake):
                assert lo % 4 == 0 and hi % 4 == 0
                covered[lo:hi] += 1
        assert bool((covered == 1).all())              # every Adam slot has exactly one owner (one-shot bucket: rank 0)


def test_wavefront_gate_configuration():
    from lstm_tensorspark_b200.ops import cuda_lstm as CL
    T, B, H = 128, 256, 1024
    v0 = (CL._wave_variant() & ~(3 << 16))             # one counter per operand k-block
    assert CL._gate_off(v0) == 512
    assert CL._gate_cfg(v0, 2, H // 64, 4, 4 * H // 64, 2, 1, B, False) == [32, 32, 8, 4, B, 1, 0]
    assert CL._gate_cfg(v0, 2, 4 * H // 64, 1, 4 * H // 64, T + 1, -1, B, True) == [128, 32, T + 1, -1, B, 0, 1]
    v1 = v0 | (1 << 16)                                # one counter per batch tile: every CTA of the tile arrives once per step
    assert CL._gate_off(v1) == 0
    assert CL._gate_cfg(v1, 2, H // 64, 4, 4 * H // 64, 2, 1, B, False) == [2, 1, 128, 64, B, 1, 0]
    assert CL._gate_cfg(v1, 2, 4 * H // 64, 1, 4 * H // 64, T + 1, -1, B, True) == [2, 1, 64 * (T + 1), -64, B, 0, 1]


def test_bucket_queue_never_releases_a_dependent_of_its_own_producer():
    from lstm_tensorspark_b200.ops import cuda_lstm as CL
    CL.AFTER_SEQ_BWD.clear()
    fired = []
    CL._big_launch_begin()                              # GEMM 1 is launched ...
    CL.queue_after_big_launch(lambda: fired.append("bucket of GEMM 1"))     # ... and its bucket becomes ready
    CL._after_big_launch()
    assert fired == []                                  # not under GEMM 1 itself (a dependent may start while its primary runs)
    CL._big_launch_begin()                              # GEMM 2
    CL._after_big_launch()
    assert fired == ["bucket of GEMM 1"]                # under the NEXT big kernel
    CL.queue_after_big_launch(lambda: fired.append("last"))
    CL._after_big_launch(flush=True)                    # end of backward: whatever is left goes in stream order
    assert fired[-1] == "last" and not CL.AFTER_SEQ_BWD


def test_device_shard_gathers_into_given_buffers():
    """`DeviceShard.next(out=...)` (the trainer hands it the captured graph's input buffers) yields the same batches as `next()`."""
    import numpy as np
    import torch
    from lstm_tensorspark_b200 import data as D
    rng = np.random.default_rng(0)
    x = rng.standard_normal((40, 3, 5)).astype(np.float32)
    y = rng.integers(0, 4, size=40).astype(np.int64)
    a = D.DeviceShard(x, y, 8, "cpu", dtype=torch.float32, shuffle=True, seed=3)
    b = D.DeviceShard(x, y, 8, "cpu", dtype=torch.float32, shuffle=True, seed=3)
    xb, yb = torch.empty(8, 3, 5), torch.empty(8, dtype=torch.int64)
    for _ in range(12):                                   # crosses two reshuffles
        xa, ya = a.next()
        xo, yo = b.next(out=(xb, yb))
        assert xo.data_ptr() == xb.data_ptr() and torch.equal(xa, xo) and torch.equal(ya, yo)
    wrong = (torch.empty(4, 3, 5), torch.empty(4, dtype=torch.int64))      # wrong batch size: falls back to fresh tensors
    xa, _ = a.next()
    xo, _ = b.next(out=wrong)
    assert torch.equal(xa, xo) and xo.data_ptr() != wrong[0].data_ptr()
