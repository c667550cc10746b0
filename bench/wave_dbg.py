import sys, os, time, json
sys.path.insert(0, "/root/repo")
import torch
from lstm_tensorspark_b200.config import Config
from lstm_tensorspark_b200.engine import TrainEngine
from lstm_tensorspark_b200 import data as Dm
from lstm_tensorspark_b200.ops import cuda_lstm as CL
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
T, B, D, C = int(os.environ.get("T", "128")), 256, 1024, 10
cfg = Config(hidden_units="1024,1024", in_features=D, seq_len=T, batch_size=B, num_classes=C, partitions=1, sync_mode="none",
             init="scaled", learn_initial_state=False, dtype="bf16", device="cuda", learning_rate=1e-3, quiet=True)
if os.environ.get("NO_TRAP"):
    CL.SEQ_VARIANT |= 1 << 20
eng = TrainEngine(cfg, 0, 1, None, batch_size=B, device=dev, dtype=torch.bfloat16)
xs, ys = Dm.synthetic_sequences(B, T, D, C, seed=1)
x = torch.as_tensor(xs).to(dev).bfloat16(); y = torch.as_tensor(ys).to(dev)
def dump(tag):
    out = {}
    for (di, t), ent in CL._WS_PAIR.items():
        for i, nm in ((0, "head"), (1, "tail")):
            ws = ent[i * 8192:(i + 1) * 8192]
            out[f"{t}_{nm}"] = {"err": int(ws[-1]), "ctr": [int(v) for v in ws[512:512 + 32 * 4:32].cpu()]}
        dn = ent[2 * 8192:]
        out[f"{t}_done"] = {"min": int(dn.min()), "max": int(dn.max()), "zeros": int((dn == 0).sum()), "n": int(dn.numel())}
    print("DBG", tag, json.dumps(out), flush=True)
for step in range(3):
    t0 = time.time()
    if os.environ.get("FWD_ONLY"):
        with torch.no_grad():
            h = eng.model.features(x)
    else:
        try:
            loss = eng.step(x, y)
        except Exception as e:
            print("EXC", repr(e)[:300], flush=True)
    torch.cuda.synchronize()
    print("step", step, "seconds", round(time.time() - t0, 3), flush=True)
    dump(f"step{step}")
