"""2+ GPUs: does a gradient bucket's fused allreduce + Adam really run NEXT TO a weight-gradient GEMM (programmatic dependent
launch) instead of after it?  Times GEMM alone, bucket alone, GEMM + bucket (PDL) and GEMM + bucket (plain stream order)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    from lstm_tensorspark_b200.models.flat import FlatParams
    from lstm_tensorspark_b200.ops.optim import FlatOptimizer
    from lstm_tensorspark_b200.ops.cuda_ext import ext
    from lstm_tensorspark_b200.parallel.fused_comm import FusedComm
    E = ext()
    comm = FusedComm(rank, world, dev, 60)
    n = 4 * 1024 * 1024 * 4            # 16.8 M params = 67 MB fp32
    p = torch.nn.Parameter(torch.zeros(n, device=dev))
    flat = FlatParams([p], [])
    comm.adopt(flat)
    opt = FlatOptimizer(flat, 1e-3, "adam")
    TB, H4, D = 32768, 4096, 1024
    dG = (torch.randn(TB, H4, device=dev) * 0.1).bfloat16()
    X = (torch.randn(TB, D, device=dev) * 0.1).bfloat16()
    gw = torch.zeros(H4, D, device=dev)
    q = n // 4

    def gemm():
        E.gemm2(dG, X, out=gw, a_mn=True, b_mn=True, out_fp32=True)

    def bucket(pdl, blocks):
        comm.launch_bucket(0, q, pdl=pdl, blocks=blocks)

    def both(pdl, blocks):
        comm.begin_grad_step(flat, opt)          # (bumps the step counter: a launch of its own, so BEFORE the GEMM)
        gemm()
        bucket(pdl, blocks)

    def timeit(fn, iters=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(); dist.barrier(device_ids=[rank]); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    out = {"world": world}
    out["gemm_ms"] = timeit(gemm)
    for blocks in (32, 64, 128):
        out[f"bucket16MB_alone_b{blocks}_ms"] = timeit(lambda: (comm.begin_grad_step(flat, opt), bucket(False, blocks)))
        out[f"gemm_then_bucket_serial_b{blocks}_ms"] = timeit(lambda: both(False, blocks))
        out[f"gemm_with_bucket_pdl_b{blocks}_ms"] = timeit(lambda: both(True, blocks))
    def chain(pdl):
        comm.begin_grad_step(flat, opt)
        gemm()
        for k in range(3):
            gemm()
            comm.launch_bucket(k * q, (k + 1) * q, pdl=pdl, blocks=64)
        gw.add_(1.0)                                   # stands in for the bias column sums
        comm.launch_bucket(3 * q, n, pdl=False, blocks=64)
    out["chain_4gemm_4buckets_serial_ms"] = timeit(lambda: chain(False))
    out["chain_4gemm_4buckets_pdl_ms"] = timeit(lambda: chain(True))
    import threading
    res = {}
    def in_thread():
        torch.cuda.set_device(rank)
        res["t"] = timeit(lambda: chain(True))
    th = threading.Thread(target=in_thread); th.start(); th.join()
    out["chain_pdl_from_other_thread_ms"] = res["t"]
    comm.begin_grad_step(flat, opt)
    out["full67MB_alone_ms"] = timeit(lambda: (comm.begin_grad_step(flat, opt), comm.launch_bucket(0, n)))
    if rank == 0:
        print("OVERLAP", json.dumps(out))
    comm.close()


if __name__ == "__main__":
    main()
