"""TrainEngine: the public "one training step" API (model + flat buffers + optimizer + cross-replica sync).

This is what ``bench.py``, ``__graft_entry__.smoke`` and the per-replica trainer call:

    eng = TrainEngine(cfg, rank, world_size, comm, batch_size=B)
    loss = eng.step(x, y)            # forward, backward, (fused allreduce +) optimizer update

One step replaces the reference's ``sess.run([train_op, loss], feed_dict=...)`` (/root/reference/src/rnn.py:264-267):
H2D feed, forward, backward, 14·L+2 ApplyAdam launches, D2H loss.  With ``cuda_graph=True`` the whole step is
captured once and replayed (launch-bound inner loops belong in CUDA graphs, not in a tracing compiler).
"""
from __future__ import annotations

from typing import Optional

import torch

from .config import Config
from .models.classifier import SequenceClassifier
from .models.recurrent.lstm import clear_weight_decay_collection, weight_decay_collection
from .ops import functional as F
from .ops.optim import FlatOptimizer
from .parallel.comm import Communicator


import os as _os

_BUCKET_PDL = _os.environ.get("LSTM_TS_BUCKET_PDL", "1") == "1"      # 0: buckets in plain stream order (no overlap) - for A/B timing


class TrainEngine:
    def __init__(self, cfg: Config, rank: int = 0, world_size: int = 1, comm: Optional[Communicator] = None,
                 batch_size: Optional[int] = None, device: Optional[torch.device] = None,
                 dtype: Optional[torch.dtype] = None, train_optimizer=None):
        self.cfg = cfg
        self.rank, self.world_size = rank, world_size
        self.comm = comm or Communicator(0, 1)
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if (cfg.device != "cpu" and torch.cuda.is_available()) \
                else torch.device("cpu")
        self.device = device
        if dtype is None:
            dtype = torch.bfloat16 if (device.type == "cuda" and cfg.dtype in ("auto", "bf16", "bfloat16")) else torch.float32
        self.dtype = dtype
        F.set_backend(cfg.backend)
        if cfg.deterministic and device.type == "cuda":
            from .ops import cuda_lstm
            cuda_lstm.SEQ_VARIANT = (cuda_lstm.SEQ_VARIANT & ~(7 << 12)) | (3 << 12)     # in-order operand stream
        clear_weight_decay_collection()
        seed = cfg.seed + (1000003 * (rank + 1) if cfg.independent_init else 0)
        gen = torch.Generator(device="cpu")
        gen.manual_seed(seed)
        self.model = SequenceClassifier(cfg, batch_size=batch_size, device="cpu", generator=gen)
        self.model.to(device)
        self.flat = self.model.build_flat()
        self.comm.adopt(self.flat)
        self.model.set_compute_dtype(dtype)
        if train_optimizer is not None:
            self.optimizer = train_optimizer(cfg.learning_rate)(self.flat)
        else:
            self.optimizer = FlatOptimizer(self.flat, cfg.learning_rate, cfg.optimizer, weight_decay=0.0)
        # K12: the optional L2 term of create_variable (/root/reference/src/models/recurrent/lstm.py:9-11) is folded into the
        # update kernel (g + wd * w over the LSTM weight / bias segment) instead of an autograd term over 67 MB of weights;
        # variables outside that segment that asked for decay (learned initial states) keep the autograd term
        self.optimizer.weight_decay = float(cfg.weight_decay or 0.0)
        self.optimizer.wd_numel = self.flat.lstm_numel
        seg_ids = {id(p) for p in self.model.rnn.averaged_parameters()}
        self._wd_in_kernel = [(v, fn, wd) for (v, fn, wd) in weight_decay_collection() if id(v) in seg_ids]
        self._wd_autograd = [(v, fn, wd) for (v, fn, wd) in weight_decay_collection() if id(v) not in seg_ids]
        self.sync_grads = cfg.sync_mode == "grad_allreduce" and world_size > 1
        self._bucket_plan = self._make_bucket_plan() if (self.sync_grads and hasattr(self.comm, "launch_bucket")
                                                         and cfg.grad_buckets) else None
        self._graph = None
        self._static = None
        self._bound = {}                        # (x ptr, y ptr, shape) -> (graph captured on that buffer, its loss tensor)
        self._bound_keepalive = []
        self.steps_done = 0

    # ---------------------------------------------------------------------------------------------------
    def _make_bucket_plan(self):
        """Gradient buckets in the order backward finishes them: [top layer (+ head)], ..., [layer 0].  A bucket = a
        contiguous element range of the flat buffer + the parameters that must have been written before it may be synced.
        (Reference counterpart: the one-shot reduceByKey over all weights, /root/reference/src/rnn.py:393-407 - here the sync
        of the upper layers hides under the backward recurrence of the layers below.)"""
        flat = self.flat
        if not flat._direct:
            return None
        off = {id(p): o for p, o in zip(flat.params, flat.offsets)}
        layers = list(self.model.rnn.layers)
        others = [p for p in flat.params[len(self.model.rnn.averaged_parameters()):]]
        others_direct = all(p.data_ptr() in flat._direct for p in others)
        end = [off[id(l.w_x)] for l in layers[1:]] + [flat.lstm_numel]        # end of each layer's segment
        plan = []
        for li in reversed(range(len(layers))):
            l = layers[li]
            # [w_h, bias] finishes with the layer's bias gradient, [w_x] already with its weight-gradient GEMM: two buckets per
            # layer, each launched under the GEMM / recurrence kernel that follows it in backward
            lo_x, lo_h, hi = off[id(l.w_x)], off[id(l.w_h)], end[li]
            need_h = [l.w_h, l.bias]
            if li == len(layers) - 1 and others_direct:
                hi = flat.padded_numel                     # head weights / bias follow the last layer in the flat buffer
                need_h = need_h + others
            plan.append({"lo": lo_x, "hi": lo_h, "need": {l.w_x.data_ptr()}})
            plan.append({"lo": lo_h, "hi": hi, "need": {p.data_ptr() for p in need_h}})
        if not others_direct:
            plan.append({"lo": flat.lstm_numel, "hi": flat.padded_numel, "need": None})    # autograd-accumulated: only final at the end
        return plan

    def _backward_with_buckets(self, loss: torch.Tensor):
        from .ops import cuda_lstm
        flat, comm, plan = self.flat, self.comm, self._bucket_plan
        comm.begin_grad_step(flat, self.optimizer)
        state = {"next": 0}

        def ready():
            # queue every leading bucket whose parameters have all been written; it is launched (PDL) right after the next big
            # backward kernel (weight-gradient GEMM / lower layer's recurrence), i.e. it runs next to that kernel
            while state["next"] < len(plan) - 1:
                b = plan[state["next"]]
                if b["need"] is None or (b["need"] & flat._stale):
                    break
                cuda_lstm.queue_after_big_launch(lambda b=b: comm.launch_bucket(b["lo"], b["hi"], pdl=_BUCKET_PDL, blocks=self.cfg.grad_bucket_blocks))
                state["next"] += 1

        cuda_lstm.HOOKS["grads_written"] = ready
        try:
            loss.backward()
        finally:
            cuda_lstm.HOOKS["grads_written"] = None
        flat.finalize_grads()
        cuda_lstm._after_big_launch(flush=True)             # queued, but no big kernel followed (generic path / last layer)
        for b in plan[state["next"]:]:
            comm.launch_bucket(b["lo"], b["hi"])

    def _step_eager(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        self.flat.zero_grad()
        loss, _logits, _correct = self.model(x, y)
        if self._wd_autograd:
            loss = loss + torch.stack([fn(v) * wd for (v, fn, wd) in self._wd_autograd]).sum()
        l2 = None
        if self._wd_in_kernel:                 # reported total loss includes the L2 value (of the weights this step used); its
            with torch.no_grad():              # gradient is applied by the update kernel
                l2 = torch.stack([fn(v) * wd for (v, fn, wd) in self._wd_in_kernel]).sum()
        if self._bucket_plan:
            self._backward_with_buckets(loss)      # backward + per-bucket fused allreduce / update, overlapped with backward
        else:
            loss.backward()
            self.flat.finalize_grads()
            if self.sync_grads:
                self.comm.grad_step_(self.flat, self.optimizer)
            else:
                self.optimizer.step()
        if l2 is not None:
            loss = loss.detach() + l2
        return loss.detach()

    def step(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """One full training step on this replica; returns the (detached, device) loss."""
        self.steps_done += 1
        if self._graph is None:
            return self._step_eager(x, y)
        bound = self._bound.get((x.data_ptr(), y.data_ptr(), tuple(x.shape)))
        if bound is not None:                   # the batch already sits in a buffer a graph was captured on: no staging copy
            g, sloss = bound
        else:
            sx, sy, sloss = self._static
            if x.data_ptr() != sx.data_ptr():       # (a loader may have gathered the batch straight into graph_inputs())
                sx.copy_(x, non_blocking=True)
            if y.data_ptr() != sy.data_ptr():
                sy.copy_(y, non_blocking=True)
            g = self._graph
        g.replay()
        self.optimizer.step_count += 1          # host mirror; the kernels use the device-resident counter
        return sloss

    def graph_inputs(self):
        """(x, y) input buffers of the captured graph, or None: a loader that assembles batches on the device can write them
        here directly (``DeviceShard.next(out=...)``) and ``step()`` then skips its staging copy."""
        return None if self._graph is None else self._static[:2]

    def maybe_average(self, force: bool = False):
        """Parameter-average sync point (reference semantics: once, at the end; or every ``sync_every`` steps)."""
        cfg = self.cfg
        if self.world_size <= 1 or cfg.sync_mode != "param_avg":
            return
        if force or (cfg.sync_every and self.steps_done % cfg.sync_every == 0):
            self.comm.average_params_(self.flat, cfg.average_scope)

    # ---------------------------------------------------------------------------------------------------
    def capture(self, x: torch.Tensor, y: torch.Tensor, warmup: int = 3, bind=()):
        """Capture fwd+bwd+update into one CUDA graph (static shapes).  Adam's bias correction is derived in-kernel from
        a device-resident step counter, so replays are exact.

        ``bind``: ``(x, y)`` pairs of LONG-LIVED device buffers that batches will be handed over in (the two staging slots of
        ``PinnedHostLoader``, fixed slices of a device-resident shard).  One more graph is captured directly on each of them,
        and ``step()`` replays it when it is given exactly that buffer - the 67 MB copy into the graph's own input buffer
        (27 us of a 4 ms step) disappears.  Any other tensor still goes through the staging copy."""
        assert self.device.type == "cuda"
        sx, sy = x.clone(), y.clone()
        # warm-up / capture run real updates: snapshot the training state and put it back, so capturing is not
        # `warmup` uncounted optimizer steps on one batch and the host / device Adam step counters stay equal
        opt = self.optimizer
        snap = {"data": self.flat.data.clone(), "step_count": opt.step_count,
                "m": None if opt.m is None else opt.m.clone(), "v": None if opt.v is None else opt.v.clone(),
                "step_dev": None if opt.step_dev is None else opt.step_dev.clone()}
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._step_eager(sx, sy)
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            sloss = self._step_eager(sx, sy)
        bound = {}
        for bx, by in bind:
            assert bx.shape == sx.shape and bx.dtype == sx.dtype and by.shape == sy.shape and bx.is_contiguous() and by.is_contiguous()
            gb = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gb):
                lb = self._step_eager(bx, by)
            bound[(bx.data_ptr(), by.data_ptr(), tuple(bx.shape))] = (gb, lb)
            self._bound_keepalive.append((bx, by))          # the graphs hold raw pointers into these buffers
        with torch.no_grad():
            self.flat.data.copy_(snap["data"])
            self.flat.refresh_shadow()
            if snap["m"] is not None:
                opt.m.copy_(snap["m"]); opt.v.copy_(snap["v"])
            if snap["step_dev"] is not None:
                opt.step_dev.copy_(snap["step_dev"])
            opt.step_count = snap["step_count"]
        self._graph, self._static, self._bound = g, (sx, sy, sloss), bound
        return g

    @torch.no_grad()
    def evaluate(self, x: torch.Tensor, y: torch.Tensor):
        h = self.model.features(x)
        logits = self.model.head(h)
        from .ops import reference as ref
        return ref.softmax_xent(logits, y), ref.accuracy(logits, y)
