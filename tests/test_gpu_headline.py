"""Numerics at the HEADLINE shape (BASELINE.json config 3: 2 x 1024 LSTM, T = 128, B = 256, bf16) through the public
TrainEngine API against the plain-PyTorch fp32 reference of the same model (ops/reference.py on the same bf16-rounded
weights and inputs): loss, h_T and every gradient, three consecutive steps (a reused arrival counter / stale tile image or a
dataflow race between the persistent kernels' CTAs would show up on step 2 or 3, bf16 drift over 128 steps in the
relative L2 error)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _engine(backend, dtype, dev, cfg_kw):
    from lstm_tensorspark_b200.config import Config
    from lstm_tensorspark_b200.engine import TrainEngine
    cfg = Config(partitions=1, sync_mode="none", init="scaled", learn_initial_state=False, device="cuda", quiet=True,
                 learning_rate=0.0, backend=backend, **cfg_kw)           # lr = 0: weights stay put, gradients stay inspectable
    return TrainEngine(cfg, 0, 1, None, batch_size=cfg.batch_size, device=dev, dtype=dtype)


def _rel_l2(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))


@pytest.mark.parametrize("hidden,T,B,D", [("1024,1024", 128, 256, 1024), ("256,256,256", 16, 256, 128)])
def test_headline_shape_matches_fp32_reference(hidden, T, B, D):
    from lstm_tensorspark_b200 import data as Dm
    from lstm_tensorspark_b200.ops import cuda_lstm, functional as F
    dev = torch.device("cuda", 0)
    C = 10
    kw = dict(hidden_units=hidden, in_features=D, seq_len=T, batch_size=B, num_classes=C)
    steps = 3
    xs, ys = Dm.synthetic_sequences(steps * B, T, D, C, seed=5)
    xs = torch.as_tensor(xs).to(dev).bfloat16()
    ys = torch.as_tensor(ys).to(dev)
    try:
        ref = _engine("torch", torch.float32, dev, kw)
        with torch.no_grad():
            ref.flat.data.copy_(ref.flat.data.bfloat16().float())          # both arms see the same bf16-representable weights
        w0 = ref.flat.data.clone()
        expect = []
        for s in range(steps):
            x, y = xs[s * B:(s + 1) * B].float(), ys[s * B:(s + 1) * B]
            loss = ref.step(x, y)
            with torch.no_grad():
                hT = ref.model.features(x)
            expect.append((float(loss), hT.clone(), ref.flat.grad.clone()))
        del ref
        torch.cuda.empty_cache()
    finally:
        F.set_backend("auto")
    eng = _engine("auto", torch.bfloat16, dev, kw)
    with torch.no_grad():
        eng.flat.data.copy_(w0)
        eng.flat.refresh_shadow()
    n0 = cuda_lstm.STATS["fast_fwd"], cuda_lstm.STATS["fast_bwd"]
    for s in range(steps):
        x, y = xs[s * B:(s + 1) * B], ys[s * B:(s + 1) * B]
        loss = eng.step(x, y)
        with torch.no_grad():
            hT = eng.model.features(x)
        torch.cuda.synchronize()
        cuda_lstm.check_kernel_errors(dev)
        l_ref, h_ref, g_ref = expect[s]
        assert abs(float(loss) - l_ref) <= 1e-2 * max(1.0, abs(l_ref)), (s, float(loss), l_ref)
        assert _rel_l2(hT, h_ref) <= 1e-2, (s, _rel_l2(hT, h_ref))
        for p, o in zip(eng.flat.params, eng.flat.offsets):
            a, b = eng.flat.grad[o:o + p.numel()], g_ref[o:o + p.numel()]
            assert _rel_l2(a, b) <= 1e-2, (s, getattr(p, "_ts_name", "?"), tuple(p.shape), _rel_l2(a, b))
    assert cuda_lstm.STATS["fast_fwd"] > n0[0] and cuda_lstm.STATS["fast_bwd"] > n0[1]


def test_weight_decay_in_the_update_kernel_matches_autograd_l2():
    """K12: create_variable's L2 term (reference lstm.py:9-11) folded into the flat update kernel == the autograd term."""
    from lstm_tensorspark_b200.config import Config
    from lstm_tensorspark_b200.engine import TrainEngine
    from lstm_tensorspark_b200 import data as Dm
    from lstm_tensorspark_b200.ops import functional as F
    dev = torch.device("cuda", 0)
    kw = dict(hidden_units="64,64", in_features=32, seq_len=4, batch_size=128, num_classes=5, partitions=1, sync_mode="none",
              init="scaled", learn_initial_state=False, device="cuda", quiet=True, learning_rate=1e-2, optimizer="sgd", weight_decay=0.1)
    x, y = Dm.synthetic_sequences(128, 4, 32, 5, seed=2)
    x, y = torch.as_tensor(x).to(dev), torch.as_tensor(y).to(dev)
    try:
        ref = TrainEngine(Config(backend="torch", **kw), 0, 1, None, batch_size=128, device=dev, dtype=torch.float32)
        w0 = ref.flat.data.clone()
        # autograd ground truth: loss + sum(wd * ||w||^2 / 2) over the LSTM variables
        ref.flat.zero_grad()
        loss, _, _ = ref.model(x, y)
        l2 = sum(0.1 * 0.5 * (p.float() ** 2).sum() for p in ref.model.rnn.averaged_parameters())
        (loss + l2).backward()
        w_expect = w0 - 1e-2 * ref.flat.grad
        total_expect = float(loss + l2)
    finally:
        F.set_backend("auto")
    eng = TrainEngine(Config(**kw), 0, 1, None, batch_size=128, device=dev, dtype=torch.float32)
    with torch.no_grad():
        eng.flat.data.copy_(w0)
    total = float(eng.step(x, y))
    assert abs(total - total_expect) < 1e-3 * max(1.0, abs(total_expect))
    assert (eng.flat.data - w_expect).abs().max() < 1e-4


def test_layer_wavefront_matches_sequential_layers():
    """The two-layer wavefront op (both recurrences co-resident, chained through the gated GEMM) against the same two layers run
    one after the other: forward states and every gradient, full-sequence loss so that dh_seq flows into the top layer too."""
    from lstm_tensorspark_b200.ops import cuda_lstm
    dev = torch.device("cuda", 0)
    torch.manual_seed(11)
    T, B, D, Ha, Hb = 12, 256, 256, 512, 256
    mk = lambda *s, sc=1.0: (torch.randn(*s, device=dev) * sc)
    x = mk(T, B, D, sc=0.5).bfloat16()
    pa = [mk(B, Ha, sc=0.1), mk(B, Ha, sc=0.1), mk(4 * Ha, D, sc=D ** -0.5), mk(4 * Ha, Ha, sc=Ha ** -0.5), mk(4 * Ha, sc=0.1)]
    pb = [mk(B, Hb, sc=0.1), mk(B, Hb, sc=0.1), mk(4 * Hb, Ha, sc=Ha ** -0.5), mk(4 * Hb, Hb, sc=Hb ** -0.5), mk(4 * Hb, sc=0.1)]
    wgt = mk(T, B, Hb)

    def run(pair):
        xa = x.clone().requires_grad_(True)
        a = [p.clone().requires_grad_(True) for p in pa]
        b = [p.clone().requires_grad_(True) for p in pb]
        if pair:
            assert cuda_lstm.wavefront_supported(xa, Ha, Hb)
            hs, hTa, cTa, hTb, cTb = cuda_lstm.lstm_pair_sequence(xa, a, b)
        else:
            hs_a, hTa, cTa = cuda_lstm.lstm_layer_sequence(xa, *a)
            hs, hTb, cTb = cuda_lstm.lstm_layer_sequence(hs_a, *b)
        loss = (hs.float() * wgt).sum() + hTa.float().sum() + cTa.float().sum() * 0.5 + hTb.float().sum() + cTb.float().sum() * 0.25
        loss.backward()
        torch.cuda.synchronize()
        cuda_lstm.check_kernel_errors(dev)
        return [hs.detach().float(), hTa.detach().float(), cTa.detach().float(), cTb.detach().float(), xa.grad.float()] + \
               [p.grad.float() for p in a + b]

    ref = run(False)
    n0 = cuda_lstm.STATS.get("wavefront_fwd", 0)
    got = run(True)
    assert cuda_lstm.STATS.get("wavefront_fwd", 0) == n0 + 1
    for i, (g, r) in enumerate(zip(got, ref)):
        assert _rel_l2(g, r) <= 5e-3, (i, tuple(r.shape), _rel_l2(g, r))


def test_graphs_bound_to_input_buffers_match_the_eager_step():
    """`TrainEngine.capture(bind=...)`: a step replayed from a graph captured directly on the buffer the batch arrives in (no
    staging copy), from the staged graph (any other tensor) and the eager step walk the same trajectory (Adam, lr > 0)."""
    from lstm_tensorspark_b200.config import Config
    from lstm_tensorspark_b200.engine import TrainEngine
    from lstm_tensorspark_b200 import data as Dm
    from lstm_tensorspark_b200.ops import cuda_lstm
    dev = torch.device("cuda", 0)
    T, B, D, C = 8, 256, 128, 10

    def mk():
        torch.manual_seed(11)
        cfg = Config(partitions=1, sync_mode="none", init="scaled", learn_initial_state=False, device="cuda", quiet=True,
                     learning_rate=1e-3, hidden_units="256,256", in_features=D, seq_len=T, batch_size=B, num_classes=C, seed=11)
        return TrainEngine(cfg, 0, 1, None, batch_size=B, device=dev, dtype=torch.bfloat16)

    xs, ys = Dm.synthetic_sequences(3 * B, T, D, C, seed=2)
    xs = torch.as_tensor(xs).to(dev).bfloat16()
    ys = torch.as_tensor(ys).to(dev)
    batches = [(xs[i * B:(i + 1) * B], ys[i * B:(i + 1) * B]) for i in range(3)]
    order = [0, 1, 2, 1, 0, 2]
    eager = mk()
    want = [float(eager.step(*batches[i])) for i in order]
    graphed = mk()
    with torch.no_grad():
        assert torch.equal(graphed.flat.data, mk().flat.data)             # same seed -> same initial weights
    graphed.capture(*batches[0], bind=batches[1:])                       # batch 0: staged path; batches 1, 2: bound graphs
    assert len(graphed._bound) == 2
    staging_before = graphed._static[0].clone()
    got = []
    for i in order:
        got.append(float(graphed.step(*batches[i])))
        if i != 0:
            assert torch.equal(graphed._static[0], staging_before)        # a bound replay never touches the staging buffer
        else:
            staging_before = graphed._static[0].clone()
    for a, b in zip(got, want):
        assert abs(a - b) <= 2e-3 * max(1.0, abs(b)), (got, want)
    assert _rel_l2(graphed.flat.data, eager.flat.data) < 1e-3
    assert graphed.optimizer.step_count == eager.optimizer.step_count == len(order)
    cuda_lstm.check_kernel_errors(dev)
