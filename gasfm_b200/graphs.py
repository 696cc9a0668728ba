"""CUDA-graph capture of a whole GASFM step (forward + loss + backward [+ gradient all-reduce]).

A step launches ~3,300 kernels (about 800 of them through the C ABI); once the per-GPU work is small
(short scenes, or a big scene sharded over 8 GPUs) the Python / launch overhead of ~80 ms per step
dominates.  Capturing the step once and replaying the graph removes it: the scene's index, the segment
plans, every workspace and every activation live at fixed addresses inside the graph's memory pool.

    step = GraphedStep(model, scene, loss_fn)        # warm-up on a side stream, then capture
    loss = step()                                    # graph replay; gradients are in p.grad

The captured scene is static: a new scene needs a new capture (``bench.py``'s end-to-end loop, which feeds
a fresh host scene every step, therefore runs eagerly).  Single-GPU only for now: capturing the NCCL
collectives of a track-sharded step deadlocked on this stack (torch 2.11 / NCCL 2.28), so ``gasfm_b200.dist``
steps run eagerly.
"""
import torch


class GraphedStep:
    def __init__(self, model, scene, loss_fn, after_backward=None, warmup=3):
        self.model, self.scene = model, scene
        self._loss_fn, self._after = loss_fn, after_backward
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager_step()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        model.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.out, self.loss = self._eager_step(zero=False)

    def _eager_step(self, zero=True):
        if zero:
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
