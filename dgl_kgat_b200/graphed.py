"""One CF iteration of the reference loop (kgat.py:139-165) as a single CUDA graph.

    step = GraphedCFStep(model, train_g, batch_size=10240)       # warms up, then captures
    loss = step(user_ids, item_pos_ids, item_neg_ids)            # host (pinned) or device int64 ids
    optimizer.step()                                             # gradients are in p.grad as usual

The captured body is exactly the public call sequence - ``compute_attention`` (optional; the
reference refreshes it once per epoch), ``g.edata['w'] = A_w``, ``gnn``, ``get_loss``,
``loss.backward()`` - so the numbers are those of the eager path; the graph only removes the
per-kernel launch latency between the ~20 kernels of a step.  Inputs travel through three static
device buffers (one async copy each per call).  Dropout draws fresh masks on every replay (torch's
graph-safe Philox state).

Capture runs on a side stream, so no autograd graph of an earlier EAGER step may still be alive when a
step is captured (drop / detach old ``loss`` and ``embedding`` tensors first): the parameters'
AccumulateGrad nodes stay bound to the stream they were created on for as long as such a graph
references them, and torch then refuses the capture ("legacy stream depend on a capturing stream").
"""
from __future__ import annotations

import torch as th

from . import _lib


class GraphedCFStep:
    def __init__(self, model, g, batch_size, refresh_attention=True, warmup=3):
        self.model, self.g, self.refresh_attention = model, g, refresh_attention
        self._cap_w = self._seen_w = None
        if not refresh_attention:
            if "w" not in g.edata:
                raise KeyError("edata['w'] is not set and refresh_attention is False")
            # The capture bakes in the ADDRESSES of the CSR / CSC weight layouts that are current now.  The
            # reference refreshes the attention once per epoch (kgat.py:139-145) and every refresh (and every
            # eval) produces new weight tensors, so the graph owns these two buffers and __call__ copies the
            # graph's current weights into them whenever they have changed since the last replay.
            dev = model.entity_embed.weight.device
            self._cap_w = self._seen_w = g.edge_weight_layouts(dev)
        self._capture(model, batch_size, warmup)

    def _capture(self, model, batch_size, warmup):
        dev = model.entity_embed.weight.device
        if dev.type != "cuda":
            raise _lib.KgatError("a graphed step needs the model on a CUDA device")
        self.device = dev
        # distinct ids for the warm-up / capture run (all-equal ids would serialise the compact-gradient
        # accumulation of the loss head on one node)
        n_rows = model.entity_embed.weight.shape[0]
        self.ids = [(th.arange(batch_size, dtype=th.long, device=dev) * (k + 1)) % n_rows for k in range(3)]
        if _lib.RECORDER.enabled and _lib.RECORDER.time_kernels:
            raise _lib.KgatError("disable per-kernel event timing before capturing a graph")
        with th.cuda.device(dev):
            side = th.cuda.Stream()
            side.wait_stream(th.cuda.current_stream())
            with th.cuda.stream(side):
                for _ in range(max(warmup, 1)):
                    self._body()
            th.cuda.current_stream().wait_stream(side)
            th.cuda.synchronize()
            for p in model.parameters():
                p.grad = None
            self.graph = th.cuda.CUDAGraph()
            with th.cuda.graph(self.graph):
                self.loss = self._body()
            # the graph writes the gradients into these static tensors on every replay
            self._grads = [(p, p.grad) for p in model.parameters() if p.grad is not None]

    def _body(self):
        m, g = self.model, self.g
        for p in m.parameters():
            p.grad = None
        if self.refresh_attention:
            with th.no_grad():
                A_w = m.compute_attention(g)
            g.edata["w"] = A_w
        emb = m.gnn(g, g.ndata["id"])
        loss = m.get_loss(emb, *self.ids)
        loss.backward()
        return loss

    def _sync_weights(self):
        """Bring the captured weight buffers up to date with ``g.edata['w']`` (two device copies, only after
        a refresh): without this a graph captured in epoch 1 would train every later epoch on epoch-1
        attention."""
        if self._cap_w is None:
            return
        cur = self.g.edge_weight_layouts(self.device)
        if cur[0] is not self._seen_w[0] or cur[1] is not self._seen_w[1]:
            if cur[0] is not self._cap_w[0]:
                self._cap_w[0].copy_(cur[0])
                self._cap_w[1].copy_(cur[1])
            self._seen_w = cur

    def __call__(self, src_ids, pos_ids, neg_ids):
        self._sync_weights()
        for dst, src in zip(self.ids, (src_ids, pos_ids, neg_ids)):
            dst.copy_(th.as_tensor(src), non_blocking=True)
        self.graph.replay()
        for p, g in self._grads:            # re-attach in case the caller cleared .grad (zero_grad(set_to_none))
            p.grad = g
        return self.loss


class GraphedDistCFStep(GraphedCFStep):
    """The same for the destination-partitioned step of :class:`dist.DistKGAT`, one CUDA graph per rank.
    With the peer exchange (default) the step consists of this library's kernels only - the row pushes
    ride on the compute kernels and the device-side barriers are ordinary kernel nodes - so the graph
    holds no NCCL work and replays without the host: 1.15 ms per step on 2 B200s (eager 1.30 ms;
    NCCL exchange, eager: 2.06 ms).  With ``exchange="nccl"`` the collectives are captured too;
    that variant is EXPERIMENTAL (gain ~3 %, process-group teardown after capturing NCCL work was
    unreliable) and bench.py uses it only when KGAT_BENCH_GRAPH_DIST=1."""

    def __init__(self, dk, batch_size, warmup=3, refresh_attention=True):
        """``refresh_attention=False``: the training loop of kgat.py, which refreshes the attention once per EPOCH
        (:139-145) and runs every CF iteration of the epoch on those weights.  The graph then holds propagation,
        loss and backward only; the weights of the own in-edges live in the arena (same address after every
        refresh), the CSC layout of the out-edge weights is a buffer this object owns and re-fills after a
        refresh (``_sync_weights``)."""
        self.dk, self.model, self.refresh_attention = dk, dk.model, refresh_attention
        self._cap_w = self._seen_w = None
        self._cap_wcsc = self._seen_wcsc = None
        if not refresh_attention:
            dk.attention_weights()                  # barrier + CSC layout of the weights that are current now
            self._cap_wcsc = self._seen_wcsc = dk._plan.w_csc
        self._capture(dk.model, batch_size, warmup)

    def _sync_weights(self):
        if self.refresh_attention:
            return
        dk = self.dk
        dk.attention_weights()                      # no-op unless compute_attention() ran since the last call
        cur = dk._plan.w_csc
        if cur is not self._seen_wcsc:
            if cur is not self._cap_wcsc:
                self._cap_wcsc.copy_(cur)
            self._seen_wcsc = cur
        dk._plan.w_csc = self._cap_wcsc

    def __call__(self, src_ids, pos_ids, neg_ids):
        self.dk.poll_peers()                # a lost peer must surface as an error, not as corrupted gradients
        return super().__call__(src_ids, pos_ids, neg_ids)

    def _body(self):
        dk = self.dk
        for p in self.model.parameters():
            p.grad = None
        if self.refresh_attention:
            dk.compute_attention()
        emb = dk.gnn()
        loss = dk.get_loss(emb, *self.ids)
        loss.backward()
        return loss


class GraphedKGStep:
    """One TransR iteration of the reference loop (kgat.py:117-131) - ``model.transR(h, r, pos_t, neg_t)``
    and ``loss.backward()`` - as a CUDA graph for a FIXED batch size (the ~40 small torch kernels of
    the TransR head are launch-bound otherwise).  Call it with device or pinned-host int64 ids; a
    batch of another size (the last, partial one) must go through the eager calls."""

    def __init__(self, model, batch_size, warmup=3):
        self.model = model
        dev = model.entity_embed.weight.device
        if dev.type != "cuda":
            raise _lib.KgatError("a graphed step needs the model on a CUDA device")
        n_ent, n_rel = model.entity_embed.weight.shape[0], model.relation_embed.weight.shape[0]
        ar = th.arange(batch_size, dtype=th.long, device=dev)
        self.batch_size = batch_size
        self.ids = [ar % n_ent, ar % n_rel, (ar * 3 + 1) % n_ent, (ar * 7 + 2) % n_ent]
        with th.cuda.device(dev):
            side = th.cuda.Stream()
            side.wait_stream(th.cuda.current_stream())
            with th.cuda.stream(side):
                for _ in range(max(warmup, 1)):
                    self._body()
            th.cuda.current_stream().wait_stream(side)
            th.cuda.synchronize()
            for p in model.parameters():
                p.grad = None
            self.graph = th.cuda.CUDAGraph()
            with th.cuda.graph(self.graph):
                self.loss = self._body()
            self._grads = [(p, p.grad) for p in model.parameters() if p.grad is not None]

    def _body(self):
        for p in self.model.parameters():
            p.grad = None
        loss = self.model.transR(*self.ids)
        loss.backward()
        return loss

    def __call__(self, h, r, pos_t, neg_t):
        src = (h, r, pos_t, neg_t)
        if all(isinstance(t, th.Tensor) and t.is_cuda and t.dtype == th.long for t in src):
            th._foreach_copy_(self.ids, list(src))                 # one multi-tensor launch
        else:
            for dst, t in zip(self.ids, src):
                dst.copy_(th.as_tensor(t), non_blocking=True)
        self.graph.replay()
        for p, g in self._grads:
            p.grad = g
        return self.loss
