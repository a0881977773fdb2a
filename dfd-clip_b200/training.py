"""Training-step driver for BASELINE config C5 (frozen encoder, trainable decoder / adapter): the body of the
reference's trainer loop — ``Detector.forward(train=True)``, ``loss.backward()``, ``optimizer.step()``
(src/trainer.py:147-178) — captured ONCE into a CUDA graph and replayed per batch.

Why: at the reference's 12 clips per GPU the native encoder needs 3.2 ms, while the one-token-per-clip decoder under
torch autograd plus the optimizer step is ~300 small launches whose cost is host dispatch, not GPU time (measured
6.1 ms per eager step on a B200). A replayed graph has no per-launch host cost.

The step computes exactly what the eager sequence computes (same kernels in the same order on the same buffers);
``tests/test_train_gpu.py::test_graphed_train_step_matches_eager_steps`` pins it against eager steps.
"""
import torch


class GraphedTrainStep:
    """``step = GraphedTrainStep(detector, optimizer, x, y, m)`` then ``loss, logits = step(x, y, m)`` per batch.

    * ``x`` fp32 (normalised) or uint8 ``[B,T,3,R,R]``, ``y`` int64 ``[B]`` (labels of task ``single_task``), ``m`` bool
      ``[B,T]`` — device tensors; every later batch must have the shapes / dtypes of the example batch (the reference
      trains with ``drop_last`` fixed-size batches, src/datasets.py loaders via main.py:246-262).
    * The example batch is used for warm-up (optimizer state creation, allocator warm-up, lazy kernel attributes); the
      parameters are restored and the optimizer state zeroed afterwards, so constructing the step does not train.
    * ``loss`` (scalar: mean task loss + auxiliary losses) and ``logits`` are views of static buffers that the next call
      overwrites.
    * Not supported (raise): ``train_mode.patch_mask`` (its patch indices are drawn on the host with numpy for every
      step, reference :511-544) and anything that needs a host decision inside the step. Gradient all-reduce for
      multi-GPU training is not part of the captured step.
    """

    def __init__(self, detector, optimizer, x, y, m, single_task=0, speed=None, warmup=2):
        if "patch_mask" in detector.train_mode:
            raise NotImplementedError("train_mode.patch_mask draws patch indices on the host for every step: "
                                      "it cannot be replayed from a CUDA graph; use the eager step")
        dev = x.device
        if dev.type != "cuda":
            raise RuntimeError("GraphedTrainStep needs CUDA tensors (dfdclip_b200 has no CPU path)")
        self.det, self.opt, self.task = detector, optimizer, int(single_task)
        self.x, self.y, self.m = x.clone(), y.clone(), m.clone()
        self.speed = None if speed is None else speed.clone()
        params = [p for g in optimizer.param_groups for p in g["params"]]
        saved = [p.detach().clone() for p in params]
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, int(warmup))):
                optimizer.zero_grad(set_to_none=True)
                self._body()
            # undo the warm-up: parameters back, optimizer state as freshly created (zero momentum / moments / step)
            with torch.no_grad():
                for p, s in zip(params, saved):
                    p.copy_(s)
                for state in optimizer.state.values():
                    for v in state.values():
                        if torch.is_tensor(v):
                            v.zero_()
            optimizer.zero_grad(set_to_none=True)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss, self.logits = self._body()

    def _body(self):
        with torch.enable_grad():
            losses, logits, other = self.det(self.x, [self.y] * (self.task + 1), self.m, speed=self.speed, train=True,
                                             single_task=self.task)
            loss = losses[self.task].mean()
            for v in other.values():
                loss = loss + v
            loss.backward()
        self.opt.step()
        return loss.detach(), logits[self.task].detach()

    def __call__(self, x, y, m, speed=None):
        if x.shape != self.x.shape or x.dtype != self.x.dtype or y.shape != self.y.shape or m.shape != self.m.shape:
            raise ValueError("batch %s/%s/%s does not match the captured step %s/%s/%s" % (
                tuple(x.shape), tuple(y.shape), tuple(m.shape), tuple(self.x.shape), tuple(self.y.shape),
                tuple(self.m.shape)))
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        self.m.copy_(m, non_blocking=True)
        if self.speed is not None:
            if speed is None:
                raise ValueError("the captured step takes a `speed` tensor")
            self.speed.copy_(speed, non_blocking=True)
        self.graph.replay()
        return self.loss, self.logits
