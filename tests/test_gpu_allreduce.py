"""Fused NVLink allreduce kernel (GPU, >= 2 devices): bit-identical replicas, exact fp32 sum order, fused SGD / Adam
epilogues vs the reference update, one-shot and two-shot, peer-pointer and multicast variants, odd message sizes."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, n_elems):
    import torch
    import torch.distributed as dist
    from lstm_tensorspark_b200.models.flat import FlatParams
    from lstm_tensorspark_b200.ops import reference as ref
    from lstm_tensorspark_b200.ops.optim import FlatOptimizer
    from lstm_tensorspark_b200.parallel.fused_comm import FusedComm
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    comm = FusedComm(rank, world, dev, 60)
    out = {}
    torch.manual_seed(7)
    base = torch.randn(n_elems)                                   # identical on every rank
    p = torch.nn.Parameter(base.clone().to(dev))
    flat = FlatParams([p], [])
    comm.adopt(flat)
    n = flat.padded_numel
    for force in ("one_shot", "two_shot"):
        # ---- parameter average --------------------------------------------------------------
        flat.data.copy_(torch.arange(n, device=dev, dtype=torch.float32) * 1e-3 + rank)
        torch.cuda.synchronize(); dist.barrier(device_ids=[rank])
        comm.average_params_(flat, "all", force=force)
        torch.cuda.synchronize()
        exp = torch.arange(n, device=dev, dtype=torch.float32) * 1e-3 + (world - 1) / 2.0
        out[f"avg_{force}"] = float(((flat.data - exp).abs() / exp.abs().clamp_min(1.0)).max())
        out[f"avg_shadow_{force}"] = float((flat.shadow.float() - flat.data).abs().max() / (flat.data.abs().max()))
        # ---- gradient allreduce + Adam ------------------------------------------------------
        flat.data.copy_(torch.linspace(-1, 1, n, device=dev))
        opt = FlatOptimizer(flat, 1e-2, "adam")
        p_ref = flat.data.clone(); m_ref = torch.zeros_like(p_ref); v_ref = torch.zeros_like(p_ref)
        for step in (1, 2):
            # bounded away from zero: Adam's update is ill-conditioned where sum(g) ~ 0, and the in-switch (NVLS) summation
            # order legitimately differs from the reference order
            gs = [(1.0 + 0.5 * torch.sin(torch.arange(n, dtype=torch.float32) * 0.01 * (r + 1) + step)).to(dev) for r in range(world)]
            flat.grad.copy_(gs[rank])
            torch.cuda.synchronize(); dist.barrier(device_ids=[rank])
            comm.grad_step_(flat, opt, force=force)
            torch.cuda.synchronize()
            gsum = gs[0].clone()
            for r in range(1, world):
                gsum += gs[r]
            ref.adam_step_(p_ref, gsum / world, m_ref, v_ref, step, 1e-2)
        # two-shot updates only this rank's slice of (m, v); the parameters are complete on every rank
        out[f"adam_{force}"] = float((flat.data - p_ref).abs().max())
        gathered = [torch.zeros_like(flat.data) for _ in range(world)]
        dist.all_gather(gathered, flat.data)
        out[f"identical_{force}"] = bool(all(torch.equal(gathered[0], t) for t in gathered))
    comm.check_errors()
    out["multicast"] = bool(comm.arena.mc_base)
    comm.close()
    return out


@pytest.mark.parametrize("n_elems", [1344, 462336 + 3])
def test_fused_allreduce(n_elems):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from lstm_tensorspark_b200.parallel.launch import launch
    world = min(torch.cuda.device_count(), 8)
    res = launch(_worker, world, args=(n_elems,))
    for r in res:
        for force in ("one_shot", "two_shot"):
            assert r[f"avg_{force}"] < 1e-6, r
            assert r[f"adam_{force}"] < 1e-5, r
            assert r[f"identical_{force}"], r
