"""Training-step driver for BASELINE config C5 (frozen encoder, trainable decoder / adapter): the body of the
reference's trainer loop — ``Detector.forward(train=True)``, ``loss.backward()``, gradient all-reduce across the ranks
(the DDP wrap of main.py:283-287 / src/trainer.py:73), ``optimizer.step()`` (src/trainer.py:147-178) — captured ONCE
into a CUDA graph and replayed per batch.

Why a graph: at the reference's 12 clips per GPU the native encoder needs ~3.2 ms while the decoder's forward /
backward is a few hundred small launches whose cost is host dispatch, not GPU time. A replayed graph has no per-launch
host cost, also with the NCCL all-reduce inside (eager DDP measured 9.3 ms per step on 8 GPUs in round 1, host bound).

Gradients live in ONE flat fp32 buffer (``.grad`` of every trainable parameter is a view into it): the native decoder
backward writes the chain's gradients straight into it (``Decoder._grad_sink``), the all-reduce is a single NCCL call
on the buffer (average, as DDP), and the fused SGD kernel reads it. The step computes what the eager sequence computes;
``tests/test_train_gpu.py`` pins it against eager steps and against the oracle's autograd.
"""
import torch


class TrainStep:
    """``step = TrainStep(detector, optimizer, x, y, m, group=None)`` then ``loss, logits = step(x, y, m)`` per batch.

    * ``x`` fp32 (normalised) or uint8 ``[B,T,3,R,R]``, ``y`` int64 ``[B]`` (labels of task ``single_task``), ``m`` bool
      ``[B,T]`` — device tensors; every later batch must have the shapes / dtypes of the example batch (the reference
      trains with fixed-size batches, main.py:246-262).
    * ``group``: a ``torch.distributed`` process group (NCCL) — every rank runs the same step on its own batch and the
      gradients are averaged across the group inside the captured step (what DDP does for the reference); None = one GPU.
    * The example batch is used for warm-up (optimizer state creation, allocator warm-up, lazy kernel attributes,
      NCCL connection set-up); parameters, module buffers and optimizer state are restored afterwards, so constructing
      the step does not train.
    * Learning-rate schedules: the step keeps ``lr`` in a device tensor that the captured optimizer kernel reads;
      call ``step.set_lr(value)`` (or ``step.sync_lr(scheduler)``) before a replay — the reference calls
      ``OneCycleLR.step()`` after every optimizer step (src/trainer.py:175-176). A momentum that changes between steps
      (``OneCycleLR(cycle_momentum=True)``) cannot be replayed and raises.
    * ``loss`` (scalar: mean task loss + auxiliary losses) and ``logits`` are views of static buffers that the next call
      overwrites.
    * ``pipeline=True``: the encoder is frozen, so its forward on batch k+1 does not depend on the optimizer step of
      batch k. The step then runs as a two-stage pipeline ACROSS calls: call k trains the decoder on batch k-1 (taps
      encoded by the previous call) while a second stream encodes batch k into the other of two tap buffers — the
      launch-bound decoder forward / backward / all-reduce / SGD hides under the tensor-bound encoder. ``step(x, y, m)``
      returns the loss / logits of the PREVIOUS batch (``(None, None)`` on the first call, which only encodes);
      ``step.flush()`` trains the last batch. The parameters after ``flush()`` equal those of the un-pipelined step on the
      same batches (every batch sees the same weights as in the reference's loop).
    * Not capturable (raise up front): ``train_mode.patch_mask`` (patch indices drawn on the host with numpy per step,
      reference :511-544), ``train_mode.temporal`` (host-side argsort / shuffle, :676-736) and class-weighted losses
      (a host list turned into a tensor inside the loss, :36-37). Use ``graph=False`` (eager body, same arithmetic).
    """

    def __init__(self, detector, optimizer, x, y, m, single_task=0, speed=None, group=None, warmup=2, graph=True,
                 pipeline=False):
        dev = x.device
        if dev.type != "cuda":
            raise RuntimeError("TrainStep needs CUDA tensors (dfdclip_b200 has no CPU path)")
        if graph:
            for key in ("patch_mask", "temporal"):
                if key in detector.train_mode:
                    raise NotImplementedError("train_mode.%s needs host work inside every step: it cannot be replayed "
                                              "from a CUDA graph; use TrainStep(..., graph=False)" % key)
            for loss in detector.config.losses:
                args = {} if isinstance(loss, str) or "args" not in loss else dict(loss.args)
                if args.get("weight"):
                    raise NotImplementedError("a class-weighted loss builds its weight tensor on the host in every "
                                              "step: not capturable; use TrainStep(..., graph=False)")
        self.det, self.opt, self.task, self.dev = detector, optimizer, int(single_task), dev
        self.group = group
        self.world = 1
        if group is not None:
            import torch.distributed as dist
            self.dist = dist
            self.world = dist.get_world_size(group)
        self.x, self.y, self.m = x.clone(), y.clone(), m.clone()
        self.speed = None if speed is None else speed.clone()
        self.pipeline = bool(pipeline)
        if self.pipeline:
            if detector.adapter is not None or speed is not None or "patch_mask" in detector.train_mode or \
                    ("ema_frame" in detector.op_mode and detector.op_mode.ema_frame):
                raise NotImplementedError("TrainStep(pipeline=True) covers the frozen-encoder step without adapter, "
                                          "patch_mask, ema_frame or auxiliary speed losses")
            enc = detector.encoder
            rows = x.shape[0] * x.shape[1] * enc.tokens_per_frame
            # two input slots and two sets of tap buffers: a call trains on one while the encoder fills the other
            self._slots = [(self.x, self.y, self.m), (x.clone(), y.clone(), m.clone())]
            self._taps = [{l: torch.empty((rows, 3 * enc.width), dtype=torch.bfloat16, device=dev)
                           for l in detector.layer_indices} for _ in range(2)]
            self._enc_stream = torch.cuda.Stream(dev)
            self._cur, self._primed = 0, False    # slot the next call trains on; whether its taps exist
        self.params = [p for g in optimizer.param_groups for p in g["params"]]
        self._build_flat_gradients()
        # lr as a device tensor: the captured fused-SGD kernel reads it at replay time
        self._lr = []
        for g in optimizer.param_groups:
            if not torch.is_tensor(g["lr"]):
                g["lr"] = torch.tensor(float(g["lr"]), dtype=torch.float32, device=dev)
            self._lr.append(g["lr"])
        self._frozen_hyper = [self._hyper(g) for g in optimizer.param_groups]
        taps, nb = detector.layer_indices, len(detector.layer_indices)
        # kernels of this library per step: encoder (stem 4, 7 per full layer, 2 for the last tap's LayerNorm + K/V
        # projection), decoder chain forward (1 + 12 per block) and backward (19 per block, 2 more between blocks, 3 at
        # the end); the tail / loss / SGD are ~30 torch kernels more
        self.launches_per_step = (4 + 7 * max(taps) + 2) + (1 + 12 * nb) + (19 * nb + 2 * (nb - 1) + 3)
        self.graph = None
        # warm-up and capture run on ONE side stream: autograd runs a parameter's gradient accumulation on the stream
        # that was current when its accumulator node was created, which must be the capturing stream. Pipelined, that
        # stream carries the decoder's small kernels beside the encoder's: high priority, so that they are placed as
        # soon as an SM frees up
        self._side = torch.cuda.Stream(dev, priority=-1) if self.pipeline else torch.cuda.Stream(dev)
        self._warm_up(max(1, int(warmup)))
        if graph:
            import gc
            gc.collect()  # no autograd graph of the warm-up may survive into the capture
            # thread_local: other threads (NVML sampling, NCCL's watchdog) may touch CUDA while this one captures
            self._graphs, self._outs = [], []
            for slot in ((0, 1) if self.pipeline else (0,)):   # pipelined: one graph per (train slot, encode slot)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self._side, capture_error_mode="thread_local"):
                    out = self._body(slot)
                self._graphs.append(g)
                self._outs.append(out)
            self.graph = self._graphs[0]
            self.loss, self.logits = self._outs[0]

    # ------------------------------------------------------------------------------------------ set-up
    @staticmethod
    def _hyper(group):
        return tuple((k, group[k]) for k in ("momentum", "weight_decay", "dampening", "nesterov", "betas", "eps")
                     if k in group)

    def _build_flat_gradients(self):
        """One fp32 buffer for all gradients; ``p.grad`` = view. Chain parameters of the decoder get their gradient
        written in place by the native backward (``_grad_sink``); the few tail parameters (ln_post, projections, ...)
        are accumulated by autograd into their (zeroed) views."""
        offs, total = [], 0
        for p in self.params:
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("TrainStep expects contiguous fp32 parameters")
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4
        self.flat = torch.zeros(total, dtype=torch.float32, device=self.dev)
        for p, o in zip(self.params, offs):
            p.grad = self.flat[o:o + p.numel()].view_as(p)
        chain = {id(p) for p in self.det.decoder._chain_params()}
        self.det.decoder._grad_sink = {id(p): p.grad for p in self.params if id(p) in chain}
        self._tail = [p for p in self.params if id(p) not in chain]
        # Several GPUs: the gradients of decoder block i are complete as soon as the native backward has walked past
        # it, so their all-reduce starts then (on a communication stream) and overlaps the backward of the blocks
        # below; what is not inside a block's contiguous slice of the flat buffer is reduced after the backward.
        self._block_slices, self._rest_slices = None, [(0, total)]
        self.det.decoder._block_grad_hook = None
        if self.world > 1:
            span = {id(p): (o, (p.numel() + 3) // 4 * 4) for p, o in zip(self.params, offs)}
            slices = []
            for blk in self.det.decoder.transformer.resblocks:
                ids = [id(p) for p in blk.parameters()]
                if not ids or any(i not in span for i in ids):
                    slices = None
                    break
                lo = min(span[i][0] for i in ids)
                hi = max(span[i][0] + span[i][1] for i in ids)
                if sum(span[i][1] for i in ids) != hi - lo:   # not contiguous in the optimizer's parameter order
                    slices = None
                    break
                slices.append((lo, hi))
            if slices and all(a[1] <= b[0] for a, b in zip(slices, slices[1:])):
                self._block_slices = slices
                rest, pos = [], 0
                for lo, hi in slices:
                    if lo > pos:
                        rest.append((pos, lo))
                    pos = hi
                if pos < total:
                    rest.append((pos, total))
                self._rest_slices = rest
                self._comm = torch.cuda.Stream(self.dev)
                self.det.decoder._block_grad_hook = self._reduce_block

    def _reduce_block(self, i):
        """Called by the native decoder backward (on autograd's thread, forward stream current) once block i's
        gradients are written: average them across the ranks on the communication stream."""
        lo, hi = self._block_slices[i]
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.dev))
        self._comm.wait_event(ev)
        with torch.cuda.stream(self._comm):
            self.dist.all_reduce(self.flat[lo:hi], op=self.dist.ReduceOp.AVG, group=self.group)

    def _warm_up(self, n):
        opt, det = self.opt, self.det
        saved = [p.detach().clone() for p in self.params]
        buffers = [(b, b.detach().clone()) for b in det.buffers()]   # e.g. BatchNorm statistics of a 768-bn adapter
        side = self._side
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            if self.pipeline:   # both tap buffers hold the example batch's taps during warm-up and capture
                self._encode(0)
                self._encode(1)
            for i in range(n):
                self._body(i % 2 if self.pipeline else 0)
            # undo the warm-up: parameters and buffers back, optimizer state as freshly created
            with torch.no_grad():
                for p, s in zip(self.params, saved):
                    p.copy_(s)
                for b, s in buffers:
                    b.copy_(s)
                for state in opt.state.values():
                    for v in state.values():
                        if torch.is_tensor(v):
                            v.zero_()
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)

    # ------------------------------------------------------------------------------------------ the step
    def _encode(self, slot):
        """Frozen encoder on the clips of input slot `slot` into its tap buffers (current stream)."""
        with torch.no_grad():
            self.det.encoder.encode(self._slots[slot][0].flatten(0, 1), keep_layers=self.det.layer_indices,
                                    qkv_into=self._taps[slot])

    def _body(self, slot=0):
        """One step. Pipelined: train on input slot `slot` (taps already there) while the encoder stream fills the
        other slot's tap buffers from its clips; the two meet again at the end of the step."""
        x, y, m, taps = self.x, self.y, self.m, None
        if self.pipeline:
            main = torch.cuda.current_stream(self.dev)
            self._enc_stream.wait_stream(main)
            with torch.cuda.stream(self._enc_stream):
                self._encode(1 - slot)
            (x, y, m), taps = self._slots[slot], self._taps[slot]
        with torch.enable_grad():
            for p in self._tail:          # autograd ACCUMULATES into existing .grad tensors
                p.grad.zero_()
            losses, logits, other = self.det(x, [y] * (self.task + 1), m, speed=self.speed, train=True,
                                             single_task=self.task, taps=taps)
            loss = losses[self.task].mean()
            for v in other.values():
                loss = loss + v
            loss.backward()
        if self.world > 1:
            # DDP's gradient averaging (NCCL over NVLink, captured in the graph): the decoder blocks' slices were
            # started from inside the backward (_reduce_block); the rest of the flat buffer follows here
            main = torch.cuda.current_stream(self.dev)
            if self._block_slices is None:
                self.dist.all_reduce(self.flat, op=self.dist.ReduceOp.AVG, group=self.group)
            else:
                self._comm.wait_stream(main)
                with torch.cuda.stream(self._comm):
                    for lo, hi in self._rest_slices:
                        self.dist.all_reduce(self.flat[lo:hi], op=self.dist.ReduceOp.AVG, group=self.group)
                main.wait_stream(self._comm)
        self.opt.step()
        if self.pipeline:
            torch.cuda.current_stream(self.dev).wait_stream(self._enc_stream)
        return loss.detach(), logits[self.task].detach()

    def _load(self, x, y, m, speed, slot=0):
        if x.shape != self.x.shape or x.dtype != self.x.dtype or y.shape != self.y.shape or m.shape != self.m.shape:
            raise ValueError("batch %s/%s/%s does not match the captured step %s/%s/%s" % (
                tuple(x.shape), tuple(y.shape), tuple(m.shape), tuple(self.x.shape), tuple(self.y.shape),
                tuple(self.m.shape)))
        sx, sy, sm = self._slots[slot] if self.pipeline else (self.x, self.y, self.m)
        sx.copy_(x, non_blocking=True)
        sy.copy_(y, non_blocking=True)
        sm.copy_(m, non_blocking=True)
        if self.speed is not None:
            if speed is None:
                raise ValueError("the captured step takes a `speed` tensor")
            self.speed.copy_(speed, non_blocking=True)
        for g, frozen in zip(self.opt.param_groups, self._frozen_hyper):
            if self.graph is not None and self._hyper(g) != frozen:
                raise RuntimeError("optimizer hyper-parameters changed since the step was captured (%s -> %s): a "
                                   "replayed graph would ignore them (e.g. OneCycleLR(cycle_momentum=True)); re-create "
                                   "the TrainStep or disable the cycling" % (dict(frozen), dict(self._hyper(g))))

    def set_lr(self, lr):
        """Set the learning rate(s) the next steps use: one value for all parameter groups or one per group."""
        values = [lr] * len(self._lr) if not isinstance(lr, (list, tuple)) else list(lr)
        for t, g, v in zip(self._lr, self.opt.param_groups, values):
            t.fill_(float(v))
            g["lr"] = t      # a scheduler may have replaced the tensor by a float

    def sync_lr(self, scheduler):
        """After ``scheduler.step()``: copy the scheduler's current learning rates into the captured step."""
        self.set_lr([float(v) for v in scheduler.get_last_lr()])

    def _run(self, slot=0):
        if self.graph is None:
            self.loss, self.logits = self._body(slot)
        else:
            self._graphs[slot].replay()
            self.loss, self.logits = self._outs[slot]
        return self.loss, self.logits

    def _push(self, x, y, m, speed=None, loaded=None):
        """One call of the step on device tensors; `loaded()` is called once the batch has been copied into the step's
        static buffers (the caller's tensors are free again). Pipelined: returns the result of the PREVIOUS batch, None
        for the first batch (which is only encoded)."""
        loaded = loaded or (lambda: None)
        if not self.pipeline:
            self._load(x, y, m, speed)
            loaded()
            return self._run()
        if not self._primed:
            self._load(x, y, m, speed, slot=self._cur)
            loaded()
            self._encode(self._cur)
            self._primed = True
            return None
        cur = self._cur
        self._load(x, y, m, speed, slot=1 - cur)
        loaded()
        out = self._run(cur)
        self._cur = 1 - cur
        return out

    def __call__(self, x, y, m, speed=None):
        out = self._push(x, y, m, speed)
        return (None, None) if out is None else out

    def flush(self):
        """Pipelined step: train the last batch handed in (its successor's encoder pass runs on stale clips and is
        discarded). Returns that batch's ``(loss, logits)``, or ``(None, None)`` if nothing is pending."""
        if not self.pipeline or not self._primed:
            return None, None
        out = self._run(self._cur)
        self._primed = False
        return out

    def run_host(self, batches):
        """The trainer loop over HOST batches (the reference's DataLoader hands out CPU tensors that the Accelerate-
        prepared loader copies to the device per step): ``batches`` yields ``(x, y, m)`` host tensors (pinned for real
        overlap); the H2D copy of batch k+1 runs on a copy stream while step k computes. Yields ``(loss, logits)`` per
        batch, views of the step's static buffers."""
        main = torch.cuda.current_stream(self.dev)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(self.dev)
            with torch.cuda.stream(self._copy_stream):   # two staging sets, owned by the copy stream's allocator pool
                self._staging = [[torch.empty_like(t) for t in (self.x, self.y, self.m)] for _ in range(2)]
            self._staged_free = [None, None]             # event: the step has copied the slot into its static buffers
        copy = self._copy_stream

        def issue(slot, batch):
            with torch.cuda.stream(copy):
                if self._staged_free[slot] is not None:
                    copy.wait_event(self._staged_free[slot])
                for dst, src in zip(self._staging[slot], batch):
                    dst.copy_(src, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy)
            return slot, ev

        it = iter(batches)
        try:
            nxt = issue(0, next(it))
        except StopIteration:
            return
        k = 0
        while nxt is not None:
            slot, ev = nxt
            k += 1
            try:
                nxt = issue(k % 2, next(it))   # in flight while this step runs
            except StopIteration:
                nxt = None
            main.wait_event(ev)

            def release(slot=slot):   # the staged batch is in the step's static buffers: the copy stream may refill it
                free = torch.cuda.Event()
                free.record(main)
                self._staged_free[slot] = free

            out = self._push(*self._staging[slot], None if self.speed is None else self.speed, loaded=release)
            if out is not None:
                yield out
        if self.pipeline and self._primed:
            yield self.flush()

    def close(self):
        """Release the captured graph and unhook from the detector. On several GPUs call this BEFORE
        ``torch.distributed.destroy_process_group()``: NCCL keeps the plans of captured collectives alive until their
        graph is destroyed, and tearing the communicator down first waits for that forever."""
        import gc
        torch.cuda.synchronize(self.dev)
        self.graph = None
        self._graphs, self._outs = [], []
        self.loss = self.logits = None
        self.det.decoder._block_grad_hook = None   # bound method: detector -> step -> detector cycle otherwise
        self.det.decoder._grad_sink = None
        gc.collect()
        torch.cuda.synchronize(self.dev)

    def eager(self, x, y, m, speed=None, serial=False):
        """The same step without the graph (per-kernel timing, debugging). Pipelined: trains on the pending batch (the
        example batch if none) and encodes this one; ``serial=True`` runs the un-pipelined body instead (encoder, then
        decoder, on one stream: what per-kernel timing needs)."""
        if not self.pipeline or serial:
            was, self.pipeline = self.pipeline, False
            try:
                self._load(x, y, m, speed)
                return self._body()
            finally:
                self.pipeline = was
        cur = self._cur
        self._load(x, y, m, speed, slot=1 - cur)
        out = self._body(cur)
        self._cur, self._primed = 1 - cur, True
        return out

    def describe(self):
        return ("Detector.forward(train=True) + backward + %sfused SGD, %s; native encoder, native decoder chain "
                "forward/backward (dfd_decoder_train_*), gradients in one flat fp32 buffer" % (
                    "all_reduce(AVG) over %d ranks (%s) + " % (
                        self.world, "per decoder block, overlapped with the backward" if self._block_slices
                        else "one call") if self.world > 1 else "",
                    ("one CUDA graph replay per step" if self.graph is not None else "eager") +
                    ("; two-stage pipeline across steps: the encoder of batch k+1 runs beside the decoder forward / "
                     "backward / optimizer step of batch k" if self.pipeline else "")))


class GraphedTrainStep(TrainStep):
    """Single-GPU ``TrainStep`` under its round-1 name."""

    def __init__(self, detector, optimizer, x, y, m, single_task=0, speed=None, warmup=2):
        super().__init__(detector, optimizer, x, y, m, single_task=single_task, speed=speed, group=None, warmup=warmup)
