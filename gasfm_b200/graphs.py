"""CUDA-graph capture of a whole GASFM step (forward + loss + backward [+ gradient exchange]).

A step launches ~3,300 kernels (about 800 of them through the C ABI); once the per-GPU work is small
(short scenes, or a big scene sharded over 8 GPUs) the Python / launch overhead of ~80 ms per step
dominates.  Capturing the step once and replaying the graph removes it: the scene's index, the segment
plans, every workspace and every activation live at fixed addresses inside the graph's memory pool.

    step = GraphedStep(model, scene, loss_fn)        # warm-up on a side stream, then capture
    loss = step()                                    # graph replay; gradients are in p.grad

Track-sharded steps are captured the same way: their cross-GPU exchanges are plain kernel launches over
peer memory (``gasfm_b200.dist.PeerExchange``), not NCCL calls, and the exchange sequence number lives in
device memory, so every rank replays its own graph and the kernels synchronise among themselves.  Pass the
``LocalGradBucket`` hooks as ``before_forward`` / ``after_backward``.  Every rank must run the same number
of warm-up steps and replays.

``GraphedStep`` captures a scene that is resident and already indexed: a new scene needs a new capture.
``StreamedStep`` captures the whole host-to-host step instead -- the CSR/CSC index build, the chunk tables and the
global plans are inside the graph -- so a NEW scene of the same shape signature (``Scene.signature()``: m, n, E, number
of valid views / tracks) is served by five async H2D copies into the captured buffers, one graph replay and the D2H
of the results, with a single host synchronisation at the end (the index's status word rides back with the results).
That covers the reference's single-scene optimisation / fine-tuning loop (same scene every step, new augmentation of
its values, ``code/single_scene_optimization.py``) and repeated inference on scenes of one shape.
"""
import torch

from . import index as _index


class GraphedStep:
    def __init__(self, model, scene, loss_fn, after_backward=None, warmup=3, before_forward=None, capture_lock=None):
        """``capture_lock``: a ``threading.Lock`` shared by ranks that live in ONE process (the single-GPU tests): their
        warm-up steps run concurrently (the exchanges need every rank), the captures one after the other."""
        self.model, self.scene = model, scene
        self._loss_fn, self._after, self._before = loss_fn, after_backward, before_forward
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager_step()
        cur.wait_stream(side)
        side.synchronize()
        cur.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        if self._before is None:
            model.zero_grad(set_to_none=True)       # gradients are (re)allocated inside the graph's pool
        torch.cuda.empty_cache()                    # the warm-up's cached blocks cannot serve the graph's private pool
        if capture_lock is not None:
            capture_lock.acquire()
        try:
            with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                self.out, self.loss = self._eager_step(zero=False)
        finally:
            if capture_lock is not None:
                capture_lock.release()

    def _eager_step(self, zero=True):
        if self._before is not None:
            self._before()
        elif zero:
            self.model.zero_grad(set_to_none=True)
        out = self.model(self.scene)
        loss = self._loss_fn(out)
        loss.backward()
        if self._after is not None:
            self._after()
        return out, loss

    def __call__(self):
        self.graph.replay()
        return self.loss


class StreamedStep:
    """Host scene in, host results out, one CUDA graph in between.

        step = StreamedStep(model, scene_host, loss_fn, outputs=("Ps_norm", "pts3D"))
        res = step(scene_host_k)        # any scene with scene_host's signature(); res["loss"], res["Ps_norm"], ... on the host

    ``scene_host`` should be pinned (``Scene.pin_memory()``) for the copies to be asynchronous.  The observation index of
    every streamed scene is validated (out-of-range / unsorted indices raise ``ValueError``) AFTER the step, when the status
    word arrives with the results; the gradients of such a step are garbage and must not be applied."""

    def __init__(self, model, scene_host, loss_fn, outputs=("Ps_norm", "pts3D"), device=None, warmup=2, train=True,
                 before_forward=None, after_backward=None):
        """``before_forward`` / ``after_backward``: the ``LocalGradBucket`` hooks of a track-sharded step (every rank streams
        its own shard; the peer-memory exchanges are captured with the rest)."""
        dev = torch.device(device) if device is not None else next(model.parameters()).device
        self.model, self._loss_fn, self._train, self._names = model, loss_fn, train, tuple(outputs)
        self._before, self._after = before_forward, after_backward
        self.scene = scene_host.to(dev)                      # the buffers the graph reads; refilled in place per step
        self.signature = scene_host.signature()
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.scene.invalidate()
                self._run()                                  # eager, validating: a malformed first scene raises here
        cur.wait_stream(side)
        side.synchronize()
        cur.synchronize()
        if self._before is None:
            model.zero_grad(set_to_none=True)                # gradients are (re)allocated inside the graph's pool
        self.scene.invalidate()
        torch.cuda.empty_cache()                             # the warm-up's cached blocks cannot serve the graph's private pool
        self.graph = torch.cuda.CUDAGraph()
        with _index.deferred_validation() as pending:
            with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                out, loss = self._run(zero=False)
                self._dev = {k: out[k].detach() for k in self._names}
                self._dev["loss"] = loss.detach().reshape(1)
                self._dev["_status"] = torch.stack([st.reshape(()) for st, _, _ in pending.entries]) if pending.entries \
                    else torch.zeros(1, dtype=torch.int32, device=dev)
        self._checks = [(m, n) for _, m, n in pending.entries]
        self._grads = [(p, p.grad) for p in model.parameters() if p.grad is not None]    # the tensors every replay rewrites
        self._host = {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for k, v in self._dev.items()}
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in self.scene._buffers().values())
        self.d2h_bytes = sum(v.numel() * v.element_size() for v in self._host.values())

    def _run(self, zero=True):
        if self._train:
            if self._before is not None:
                self._before()
            elif zero:
                self.model.zero_grad(set_to_none=True)
            out = self.model(self.scene)
            loss = self._loss_fn(out)
            loss.backward()
            if self._after is not None:
                self._after()
        else:
            with torch.no_grad():
                out = self.model(self.scene)
                loss = self._loss_fn(out)
        return out, loss

    def __call__(self, scene_host):
        self.scene.copy_from(scene_host, non_blocking=True)          # raises ValueError on a different signature
        self.graph.replay()
        for p, g in self._grads:
            p.grad = g               # (an optimizer's zero_grad(set_to_none=True) in between must not detach them)
        for k, v in self._dev.items():
            self._host[k].copy_(v, non_blocking=True)
        torch.cuda.current_stream(self._dev["loss"].device).synchronize()
        for st, (m, n) in zip(self._host["_status"].tolist(), self._checks):
            _index.raise_for_status(int(st), m, n)
        res = {k: self._host[k] for k in self._names}
        res["loss"] = float(self._host["loss"][0])
        return res
