"""Kernel timeline of ONE training step (torch.profiler / CUPTI): name, start, duration - who overlaps whom.
    python -m torch.distributed.run --nproc-per-node N bench/trace_step.py     (or plain python for one GPU)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    from lstm_tensorspark_b200.config import Config
    from lstm_tensorspark_b200.engine import TrainEngine
    from lstm_tensorspark_b200.parallel.comm import make_communicator
    from lstm_tensorspark_b200 import data as Dm
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29577")
    B, T, D, C = 256, 128, 1024, 10
    cfg = Config(hidden_units="1024,1024", in_features=D, seq_len=T, batch_size=B, num_classes=C, partitions=world,
                 sync_mode="grad_allreduce" if world > 1 else "none", average_scope="all", init="scaled", learn_initial_state=False,
                 comm="fused", dtype="bf16", device="cuda", learning_rate=1e-3, quiet=True)
    comm = make_communicator("fused" if world > 1 else "auto", rank, world, dev)
    eng = TrainEngine(cfg, rank, world, comm, batch_size=B, device=dev, dtype=torch.bfloat16)
    xs, ys = Dm.synthetic_sequences(B, T, D, C, seed=rank)
    x = torch.as_tensor(xs).to(dev).bfloat16(); y = torch.as_tensor(ys).to(dev)
    for _ in range(5):
        eng.step(x, y)
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        eng.step(x, y)
        torch.cuda.synchronize()
    if rank == 0:
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
        evs.sort(key=lambda e: e.time_range.start)
        t0 = evs[0].time_range.start
        rows = []
        for e in evs:
            nm = e.name
            for k in ("lstm_seq_kernel", "gemm2_kernel", "ar_two_shot", "ar_one_shot", "colsum", "head_", "flat_adam", "transpose", "seq_prologue"):
                if k in nm:
                    nm = k + ("<bwd>" if "ILb1E" in e.name and k == "lstm_seq_kernel" else "")
                    break
            rows.append((round(e.time_range.start - t0, 1), round(e.time_range.end - e.time_range.start, 1), nm[:60]))
        out = os.path.join(ROOT, "gpurun_out", f"trace_step_n{world}.txt")
        os.makedirs(os.path.dirname(out), exist_ok=True)
        with open(out, "w") as f:
            f.write("start_us dur_us kernel\n")
            for r in rows:
                f.write(f"{r[0]:10.1f} {r[1]:9.1f} {r[2]}\n")
        print("TRACE rows", len(rows), "span_us", rows[-1][0] + rows[-1][1])
    comm.close()


if __name__ == "__main__":
    main()
