"""Driver contract checks that need no GPU: the reference arm answers with a JSON line, the extension builds/imports."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_reports_unavailable():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and "unavailable" in d and len(d["unavailable"]) > 10


def test_extension_is_built_and_importable():
    so = [f for f in os.listdir(os.path.join(ROOT, "lstm_tensorspark_b200")) if f.startswith("_C") and f.endswith(".so")]
    if not so:                                   # fresh checkout: build it (nvcc cross-compiles without a GPU)
        sys.path.insert(0, ROOT)
        import __graft_entry__ as g
        g.build()
    from lstm_tensorspark_b200.ops.cuda_ext import ext
    E = ext()
    assert E.ar_flag_words() == E.ar_max_blocks() * 16 * E.ar_slots()
    for name in ("gemm2", "gemm_generic", "lstm_seq_fwd", "lstm_seq_bwd", "fused_allreduce", "head_fwd", "head_bwd", "flat_adam", "lstm_pointwise_fwd"):
        assert hasattr(E._m, name)


def test_cuda_op_on_cpu_tensor_is_rejected_when_forced():
    import pytest
    import torch
    from lstm_tensorspark_b200.ops import functional as F
    F.set_backend("cuda_ext")
    try:
        with pytest.raises(RuntimeError):
            F.lstm_cell_step(torch.zeros(2, 3), torch.zeros(2, 4), torch.zeros(2, 4), torch.zeros(16, 3), torch.zeros(16, 4), torch.zeros(16))
    finally:
        F.set_backend("auto")


def test_clock_sampler_degrades_without_a_gpu():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cs = mod.ClockSampler(0)
    cs.start()
    cs.mark()
    out = cs.stop()
    assert set(out) >= {"sm_mhz", "sm_max_mhz", "reasons"}
