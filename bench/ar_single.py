"""One-rank run of the fused allreduce kernels (world = 1: the cross-rank barrier and the peer / multicast addressing are
exercised against this GPU's own buffers) - the ONLY way to put them under ncu, which must never wrap a multi-rank command."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch


def main():
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29590")
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    from lstm_tensorspark_b200.models.flat import FlatParams
    from lstm_tensorspark_b200.ops.optim import FlatOptimizer
    from lstm_tensorspark_b200.parallel.fused_comm import FusedComm
    comm = FusedComm(0, 1, dev, 60)
    n = 16 * 1024 * 1024
    p = torch.nn.Parameter(torch.randn(n, device=dev))
    flat = FlatParams([p], [])
    comm.adopt(flat)
    opt = FlatOptimizer(flat, 1e-3, "adam")
    flat.grad.normal_()
    for force in ("two_shot", "one_shot"):
        for _ in range(2):
            comm.grad_step_(flat, opt, force=force)
            comm.average_params_(flat, "all", force=force)
    torch.cuda.synchronize()
    comm.check_errors()
    print("ar_single ok, multicast:", bool(comm.arena.mc_base))
    comm.close()


if __name__ == "__main__":
    main()
