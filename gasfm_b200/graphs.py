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

The captured scene is static: a new scene needs a new capture.
"""
import torch


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
