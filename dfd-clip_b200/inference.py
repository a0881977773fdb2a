"""Video-level scoring driver: the caller loop of the reference's ``inference.py:107-156`` around the hot path.

* ``HostClipStream`` / ``predict_stream``: batch after batch of host clips in, host logits out, pipelined ACROSS
  batches on three streams (H2D of batch k+1, encoder of batch k, decoder + D2H of batch k-1); the throughput path and
  what ``bench.py`` reports as ``e2e``.
* ``HostClipPipeline`` / ``predict_from_host``: one blocking call per batch (latency path): the encoder runs chunk by
  chunk with the H2D copies double-buffered on a copy stream, the decoder once per batch (the reference does a
  blocking ``.to(device)`` / ``.to("cpu")`` around every chunk, inference.py:116-118).
* ``pack_clip_batches``: the ``torch.stack(clips[i:i + N])`` of inference.py:117 as a threaded packer into rotating
  pinned staging buffers, across video boundaries.
* ``shard_videos`` / ``score_videos`` / ``score_videos_batched``: videos are independent units, sharded across ranks
  by clip count; each rank scores its videos (softmax per clip, mean over the clips of a video, inference.py:121,140;
  ``video_mean_probs`` is a deterministic fp64 segment mean) and ONE all_gather of the per-video scores assembles the
  result on every rank (the reference gathers once per video, :147-149). No collective runs inside the clip loop.
"""
import torch


def chunk_schedule(n, chunk_clips=32, first=8, second=24):
    """Clip counts of the encoder chunks for a batch of n clips: a small first chunk so the GPU starts after a short
    exposed H2D copy, then growing chunks (each chunk's copy hides behind the previous chunk's encoder)."""
    sizes, left = [], int(n)
    for want in (first, second):
        if left <= 0:
            break
        take = min(want, left)
        sizes.append(take)
        left -= take
    while left > 0:
        take = min(chunk_clips, left)
        sizes.append(take)
        left -= take
    return sizes


class HostClipPipeline:
    """Host -> device -> host pipeline around the hot path: the ENCODER runs chunk by chunk while the next chunk's
    frames are copied on a second stream (double-buffered); every chunk writes its K/V taps into one set of
    whole-batch tap buffers; the DECODER (one query per clip, launch-bound at small batches) then runs ONCE over the
    whole batch. The reference does a blocking ``.to(device)`` / ``.to("cpu")`` around every chunk's full predict
    (inference.py:115-118)."""

    MAX_BATCH = 128  # clips per decoder pass (tap buffers: 6 x 930 MB at ViT-B/16, T=8)

    def __init__(self, detector, chunk_clips=32):
        self.det = detector
        self.chunk = int(chunk_clips)
        self.chunk_u8 = (16, 48, 64)  # first, second, following chunk sizes for uint8 clips
        self.dev = next(detector.decoder.parameters()).device
        if self.dev.type != "cuda":
            raise RuntimeError("HostClipPipeline needs the detector on a CUDA device")
        self.copy_stream = torch.cuda.Stream(self.dev)
        self._bufs = None
        self._taps = None

    def _buffers(self, x_host, rows):
        shape = (rows,) + tuple(x_host.shape[1:])
        if self._bufs is None or self._bufs[0].shape[1:] != shape[1:] or self._bufs[0].shape[0] < rows \
                or self._bufs[0].dtype != x_host.dtype:
            self._bufs = [torch.empty(shape, dtype=x_host.dtype, device=self.dev) for _ in range(2)]  # fp32 or uint8
        return self._bufs

    def _tap_buffers(self, n_frames):
        enc = self.det.encoder
        rows, cols = n_frames * enc.tokens_per_frame, 3 * enc.width
        if self._taps is None or next(iter(self._taps.values())).shape[0] < rows:
            self._taps = None
            self._taps = {l: torch.empty((rows, cols), dtype=torch.bfloat16, device=self.dev)
                          for l in self.det.layer_indices}
        return self._taps

    def _run_batch(self, x_host, m_host, out_host):
        n, t = x_host.shape[:2]
        det, dev, main = self.det, self.dev, torch.cuda.current_stream(self.dev)
        if x_host.dtype == torch.uint8:
            # 1 byte per pixel: the copies are four times shorter, so fewer and larger chunks win
            sizes = chunk_schedule(n, self.chunk_u8[2], first=self.chunk_u8[0], second=self.chunk_u8[1])
        else:
            sizes = chunk_schedule(n, self.chunk)
        bufs = self._buffers(x_host, max(sizes))
        taps = self._tap_buffers(n * t)
        m_dev = m_host.to(dev, non_blocking=True)
        copied = [torch.cuda.Event() for _ in sizes]
        consumed = [torch.cuda.Event() for _ in sizes]
        self.copy_stream.wait_stream(main)
        s = 0
        for i, size in enumerate(sizes):
            xb = bufs[i % 2]
            with torch.cuda.stream(self.copy_stream):
                if i >= 2:
                    self.copy_stream.wait_event(consumed[i - 2])  # buffer free again
                xb[:size].copy_(x_host[s:s + size], non_blocking=True)
                copied[i].record(self.copy_stream)
            main.wait_event(copied[i])
            det.encoder.encode(xb[:size].flatten(0, 1), keep_layers=det.layer_indices, qkv_into=taps,
                               frame_offset=s * t)
            consumed[i].record(main)
            s += size
        with torch.no_grad():
            logits, _ = det.predict_from_taps(taps, m_dev, n, t)
        out_host.copy_(logits[0], non_blocking=True)

    def __call__(self, x_host, m_host, out_host=None):
        """x_host fp32 (normalised) or uint8 (raw pixels) [B,T,3,R,R], m_host bool [B,T] (pinned for real overlap). Returns logits of task 0 on the host
        (fp32 [B, out_dim], pinned); the call returns after the last D2H copy has completed."""
        n = x_host.shape[0]
        out_dim = self.det.out_dim[0]
        if out_host is None:
            out_host = torch.empty((n, out_dim), dtype=torch.float32).pin_memory()
        for s in range(0, n, self.MAX_BATCH):
            e = min(s + self.MAX_BATCH, n)
            self._run_batch(x_host[s:e], m_host[s:e], out_host[s:e])
        torch.cuda.current_stream(self.dev).synchronize()
        return out_host


class HostClipStream:
    """Batch-after-batch scoring of host-resident clips, pipelined ACROSS batches (the reference's caller loop,
    inference.py:107-156, issues a blocking ``.to(device)``, a full ``predict`` and a blocking ``.to("cpu")`` per
    chunk). Three streams work on up to three batches at once:

    * copy stream:    H2D of batch k+1 (one whole-batch copy into the free one of two device input buffers),
    * main stream:    encoder of batch k (one full-size pass: no chunking, the GEMMs keep their full M),
    * decoder stream: decoder + logit normalisation of batch k-1 and the D2H copy of its logits
      (``overlap_decoder``; the decoder is one query token per clip, launch/HBM-bound, and reads its own of two sets
      of tap buffers, so it can run beside the next batch's tensor-bound encoder).

    ``run`` is a generator: it yields one host tensor of task-0 logits per input batch, in order, one batch behind the
    batch being issued. Results are bit-identical to ``Detector.predict`` on the same clips (same kernels, same
    launch shapes)."""

    def __init__(self, detector, overlap_decoder=True):
        self.det = detector
        self.dev = next(detector.decoder.parameters()).device
        if self.dev.type != "cuda":
            raise RuntimeError("HostClipStream needs the detector on a CUDA device")
        self.copy_stream = torch.cuda.Stream(self.dev)
        self.dec_stream = torch.cuda.Stream(self.dev, priority=-1) if overlap_decoder else None
        self._x = [None, None]     # device input buffers (fp32 or uint8)
        self._m = [None, None]     # device masks
        self._taps = [None, None]  # per-slot tap buffers {layer: bf16 [rows, 3D]}
        self._out = [None, None]   # pinned host logits

    def _slot_buffers(self, slot, n, x_tail, x_dtype, m_tail, m_dtype):
        """Device input / mask buffers, tap buffers and pinned output of one slot for a batch of ``n`` clips of shape
        ``x_tail`` (= [T, 3, R, R]) and masks of shape ``m_tail`` (= [T]); grown when a batch needs more."""
        t = x_tail[0]
        # The copy stream is the first writer of the input buffers, so they come from ITS allocator pool: a block
        # handed out on the main stream may still be written by main-stream kernels that were launched (and whose
        # tensors were freed) earlier, which the copy stream does not wait for.
        with torch.cuda.stream(self.copy_stream):
            xb = self._x[slot]
            if xb is None or xb.dtype != x_dtype or tuple(xb.shape[1:]) != tuple(x_tail) or xb.shape[0] < n:
                self._x[slot] = None
                xb = self._x[slot] = torch.empty((n,) + tuple(x_tail), dtype=x_dtype, device=self.dev)
            mb = self._m[slot]
            if mb is None or tuple(mb.shape[1:]) != tuple(m_tail) or mb.shape[0] < n or mb.dtype != m_dtype:
                mb = self._m[slot] = torch.empty((n,) + tuple(m_tail), dtype=m_dtype, device=self.dev)
        enc = self.det.encoder
        rows, cols = n * t * enc.tokens_per_frame, 3 * enc.width
        tap_slot = slot if self.dec_stream is not None else 0  # one stream -> the decoder is done before the next encoder
        taps = self._taps[tap_slot]
        if taps is None or next(iter(taps.values())).shape[0] < rows:
            self._taps[tap_slot] = None
            taps = self._taps[tap_slot] = {l: torch.empty((rows, cols), dtype=torch.bfloat16, device=self.dev)
                                           for l in self.det.layer_indices}
        ob = self._out[slot]
        out_dim = self.det.out_dim[0]
        if ob is None or ob.shape[0] < n:
            ob = self._out[slot] = torch.empty((n, out_dim), dtype=torch.float32).pin_memory()
        return xb, mb, taps, ob

    def _issue(self, slot, x_host, m_host):
        """Enqueue copy, encoder and decoder of one batch; returns (event after the D2H copy, pinned logits view).
        ``x_host`` / ``m_host`` are tensors, or equally long lists of tensors (pieces of pinned host memory that are
        copied back to back into the device batch buffer: no host-side staging pass)."""
        x_parts = list(x_host) if isinstance(x_host, (list, tuple)) else [x_host]
        m_parts = list(m_host) if isinstance(m_host, (list, tuple)) else [m_host]
        n, t = sum(int(p.shape[0]) for p in x_parts), x_parts[0].shape[1]
        det, main = self.det, torch.cuda.current_stream(self.dev)
        xb, mb, taps, ob = self._slot_buffers(slot, n, tuple(x_parts[0].shape[1:]), x_parts[0].dtype,
                                              tuple(m_parts[0].shape[1:]), m_parts[0].dtype)
        copied, done = torch.cuda.Event(), torch.cuda.Event()
        # the slot's buffers are free: the batch that used them (two back) was synchronised before this call
        with torch.cuda.stream(self.copy_stream):
            off = 0
            for xp, mp in zip(x_parts, m_parts):
                k = int(xp.shape[0])
                xb[off:off + k].copy_(xp, non_blocking=True)
                mb[off:off + k].copy_(mp, non_blocking=True)
                off += k
            copied.record(self.copy_stream)
        main.wait_event(copied)
        det.encoder.encode(xb[:n].flatten(0, 1), keep_layers=det.layer_indices, qkv_into=taps)
        tail = main
        if self.dec_stream is not None:
            encoded = torch.cuda.Event()
            encoded.record(main)
            self.dec_stream.wait_event(encoded)
            tail = self.dec_stream
        with torch.cuda.stream(tail), torch.no_grad():
            logits, _ = det.predict_from_taps(taps, mb[:n], n, t)
            ob[:n].copy_(logits[0], non_blocking=True)
            done.record(tail)
        return done, ob[:n]

    def run(self, batches):
        """``batches``: iterable of ``(x_host, m_host)`` (x fp32 normalised or uint8 ``[B,T,3,R,R]``, m bool ``[B,T]``,
        pinned for real overlap). Yields fp32 ``[B, out_dim]`` host tensors (task 0), one per batch, in order."""
        pending = None
        out_dim = self.det.out_dim[0]
        for k, (x_host, m_host) in enumerate(batches):
            empty = (sum(int(p.shape[0]) for p in x_host) == 0) if isinstance(x_host, (list, tuple)) \
                else x_host.shape[0] == 0
            if empty:
                cur = (None, torch.empty((0, out_dim), dtype=torch.float32))
            else:
                cur = self._issue(k % 2, x_host, m_host)
            if pending is not None:
                if pending[0] is not None:
                    pending[0].synchronize()
                yield pending[1].clone()
            pending = cur
        if pending is not None:
            if pending[0] is not None:
                pending[0].synchronize()
            yield pending[1].clone()


def predict_stream(detector, batches, overlap_decoder=True):
    """Generator over ``HostClipStream.run``: host batches in, host logits out, H2D / encoder / decoder+D2H of
    consecutive batches overlapped."""
    yield from HostClipStream(detector, overlap_decoder=overlap_decoder).run(batches)


_PIPELINES = {}


def predict_from_host(detector, x_host, m_host, chunk_clips=32):
    """End-to-end public call: host clips in, host logits out (H2D and D2H inside)."""
    key = (id(detector), chunk_clips)
    pipe = _PIPELINES.get(key)
    if pipe is None:
        pipe = _PIPELINES[key] = HostClipPipeline(detector, chunk_clips)
    return pipe(x_host, m_host)


def shard_videos(clip_counts, world_size):
    """Deterministic balanced partition of videos over ranks by clip count (longest-processing-time greedy).
    Returns ``world_size`` lists of video indices, each ascending."""
    order = sorted(range(len(clip_counts)), key=lambda i: (-int(clip_counts[i]), i))
    loads = [0] * world_size
    shards = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += int(clip_counts[i])
    return [sorted(s) for s in shards]


def video_mean_probs(clip_logits, clip_counts):
    """softmax per clip, then mean over each video's clips (inference.py:121,140). clip_logits [sum(counts), O].
    The per-video sums are accumulated in fp64 over a padded [videos, max_count] gather — no atomics, so the scores are
    run-to-run deterministic and do not depend on how many videos are reduced in one call (index_add_ on CUDA is
    neither)."""
    probs = clip_logits.float().softmax(dim=-1)
    counts = torch.as_tensor(list(clip_counts), device=probs.device, dtype=torch.long)
    n_videos = counts.numel()
    if n_videos == 0:
        return probs.new_zeros((0, probs.shape[1]))
    max_count = max(1, int(max(clip_counts)))
    starts = torch.cumsum(counts, 0) - counts
    slot = torch.arange(max_count, device=probs.device)
    valid = slot.unsqueeze(0) < counts.unsqueeze(1)                                   # [V, max_count]
    index = (starts.unsqueeze(1) + slot.unsqueeze(0)).clamp_(max=max(probs.shape[0] - 1, 0))
    if probs.shape[0] == 0:
        return probs.new_zeros((n_videos, probs.shape[1]))
    padded = probs.double()[index] * valid.unsqueeze(-1)                              # [V, max_count, O]
    sums = padded.sum(dim=1)
    return (sums / counts.clamp_min(1).unsqueeze(1)).to(probs.dtype)


_STAGING = {}  # (shape, dtype, pinned) -> idle staging buffers of pack_clip_batches


def _take_staging(key):
    idle = _STAGING.get(key)
    if idle:
        return idle.pop()
    shape, dtype, pin = key
    return torch.empty(shape, dtype=dtype, pin_memory=pin)


def pack_clip_batches(videos, masks, batch_clips=64, pin=None, workers=4):
    """Generator: the clips of ``videos`` (host tensors ``[n_i, T, 3, R, R]``, any mix of lengths, zero allowed) packed
    in order into batches of ``batch_clips`` clips, video boundaries ignored. Each batch is a view of one of THREE
    rotating staging buffers (pinned when CUDA is available, so the H2D copy of ``HostClipStream`` is asynchronous);
    while the consumer works on batch k, a small thread pool already copies batch k+1 into the next buffer (a batch of
    64 fp32 clips is 308 MB of memcpy: longer than the GPU needs for the previous batch if done by one thread in the
    consumer's loop). A consumer may hold batch k until it asks for batch k+2, which is exactly what
    ``HostClipStream.run`` does (batch k-2 has been synchronised before batch k is requested). Nothing is
    concatenated or pinned up front: host memory beyond the caller's videos is three batches. Videos that are ALREADY
    pinned are not staged at all: their batches are yielded as lists of clip ranges (views), which
    ``HostClipStream`` copies piece by piece into the device batch buffer."""
    from concurrent.futures import ThreadPoolExecutor
    step = max(1, int(batch_clips))
    first = next((v for v in videos if v.shape[0] > 0), None)
    if first is None:
        return
    if pin is None:
        pin = torch.cuda.is_available()
    m_first = next(m for v, m in zip(videos, masks) if v.shape[0] > 0)
    # plan: per batch, the list of (video index, first clip, clip count, destination row)
    plan, cur, fill = [], [], 0
    for i, (v, m) in enumerate(zip(videos, masks)):
        if v.shape[1:] != first.shape[1:] or v.dtype != first.dtype:
            raise ValueError("all videos must share one clip shape and dtype: %s %s vs %s %s" %
                             (tuple(v.shape[1:]), v.dtype, tuple(first.shape[1:]), first.dtype))
        if m.shape[0] != v.shape[0]:
            raise ValueError("mask rows %d != clips %d" % (m.shape[0], v.shape[0]))
        a, n = 0, int(v.shape[0])
        while a < n:
            take = min(step - fill, n - a)
            cur.append((i, a, take, fill))
            fill += take
            a += take
            if fill == step:
                plan.append((cur, fill))
                cur, fill = [], 0
    if fill:
        plan.append((cur, fill))
    if pin and all(v.is_pinned() and m.is_pinned() for v, m in zip(videos, masks) if v.shape[0] > 0):
        # the caller's videos are already page-locked (e.g. a DataLoader with pin_memory=True): no staging pass at
        # all — a batch is the list of clip ranges it consists of, and HostClipStream copies them back to back
        # into its device batch buffer
        for pieces, rows in plan:
            yield ([videos[i][a:a + c] for i, a, c, _ in pieces], [masks[i][a:a + c] for i, a, c, _ in pieces])
        return
    n_slots = min(3, len(plan))
    # staging buffers come from (and go back to) a small process-wide pool: page-locking 3 x 77 MB costs ~0.1 s, as
    # much as scoring 400 clips
    key_x = ((step,) + tuple(first.shape[1:]), first.dtype, bool(pin))
    key_m = ((step,) + tuple(m_first.shape[1:]), m_first.dtype, bool(pin))
    reused = bool(_STAGING.get(key_x)) or bool(_STAGING.get(key_m))
    stage_x = [_take_staging(key_x) for _ in range(n_slots)]
    stage_m = [_take_staging(key_m) for _ in range(n_slots)]
    if reused and pin and torch.cuda.is_available():
        # a previous consumer may have queued its last H2D copies out of these buffers just before handing them back
        torch.cuda.synchronize()
    workers = max(1, int(workers))
    rows_per_task = max(1, -(-step // workers))

    def copy_rows(slot, vid, src, dst, count):
        stage_x[slot][dst:dst + count].copy_(videos[vid][src:src + count])
        stage_m[slot][dst:dst + count].copy_(masks[vid][src:src + count])

    pool = ThreadPoolExecutor(max_workers=workers)

    def fill_async(j):
        futures = []
        for vid, src, count, dst in plan[j][0]:
            for off in range(0, count, rows_per_task):
                c = min(rows_per_task, count - off)
                futures.append(pool.submit(copy_rows, j % n_slots, vid, src + off, dst + off, c))
        return futures

    try:
        pending = fill_async(0)
        for j in range(len(plan)):
            for f in pending:
                f.result()
            pending = fill_async(j + 1) if j + 1 < len(plan) else []
            rows = plan[j][1]
            yield stage_x[j % n_slots][:rows], stage_m[j % n_slots][:rows]
    finally:
        pool.shutdown(wait=True)
        for buf in stage_x:
            _STAGING.setdefault(key_x, []).append(buf)
        for buf in stage_m:
            _STAGING.setdefault(key_m, []).append(buf)


def score_videos_batched(detector, videos, masks, batch_clips=64, group=None):
    """The same result as ``score_videos`` for host-resident videos, computed the way the hardware likes it: the
    clips of ALL of this rank's videos are packed into batches of ``batch_clips`` (video boundaries do not matter:
    clips are independent) by ``pack_clip_batches`` (two rotating pinned staging buffers) and pushed through
    ``HostClipStream`` (packing + H2D of the next batch, encoder of this one and decoder + D2H of the previous one in
    flight together); the per-video mean of clip probabilities is a segment mean at the end and ONE
    all_gather assembles the ranks' scores. The reference scores one video at a time in chunks of 16 clips with a
    blocking copy in each direction per chunk (inference.py:107-156).

    ``videos[i]``: host tensor [n_i, T, 3, R, R] (fp32 normalised or uint8), ``masks[i]``: [n_i, T].
    Returns per-video mean probabilities [len(videos), O] on the detector's device, identical on every rank."""
    import torch.distributed as dist
    distributed = dist.is_available() and dist.is_initialized() and (dist.get_world_size(group) > 1)
    world = dist.get_world_size(group) if distributed else 1
    rank = dist.get_rank(group) if distributed else 0
    counts = [int(v.shape[0]) for v in videos]
    shards = shard_videos(counts, world)
    mine = [i for i in shards[rank] if counts[i] > 0]
    dev = next(detector.decoder.parameters()).device
    out_dim = detector.out_dim[0]
    result = torch.full((len(videos), out_dim), float("nan"), device=dev)
    if mine:
        pipe = HostClipStream(detector)
        batches = pack_clip_batches([videos[i] for i in mine], [masks[i] for i in mine], batch_clips)
        logits = torch.cat(list(pipe.run(batches))).to(dev)
        rows = video_mean_probs(logits, [counts[i] for i in mine])
    else:
        rows = torch.empty((0, out_dim), device=dev)
    if not distributed:
        if mine:
            result[torch.as_tensor(mine, dtype=torch.long, device=dev)] = rows
        return result
    kept = [[i for i in s if counts[i] > 0] for s in shards]
    width = max(1, max(len(k) for k in kept))
    padded = torch.full((width, out_dim), float("nan"), device=dev)
    padded[:rows.shape[0]] = rows
    gathered = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(gathered, padded, group=group)  # the single collective of the path
    for r, idx in enumerate(kept):
        if idx:
            result[torch.as_tensor(idx, dtype=torch.long, device=dev)] = gathered[r][:len(idx)]
    return result


def score_videos(predict_fn, videos, masks, chunk_clips=16, group=None, device=None):
    """Score a list of videos. ``videos[i]``: tensor [n_i, T, 3, R, R] of that video's clips, ``masks[i]``: [n_i, T].
    ``predict_fn(x, m) -> logits [n, O]`` (e.g. ``lambda x, m: det.predict(x, m)[0][0]``).
    Returns per-video mean probabilities [len(videos), O], identical on every rank. Videos with zero clips get NaN
    rows (the reference skips them, inference.py:109-111)."""
    import torch.distributed as dist
    distributed = dist.is_available() and dist.is_initialized() and (dist.get_world_size(group) > 1)
    world = dist.get_world_size(group) if distributed else 1
    rank = dist.get_rank(group) if distributed else 0
    counts = [int(v.shape[0]) for v in videos]
    mine = shard_videos(counts, world)[rank]
    local = []
    out_dim = None
    for i in mine:
        if counts[i] == 0:
            local.append(None)
            continue
        chunks = []
        for s in range(0, counts[i], chunk_clips):
            x, m = videos[i][s:s + chunk_clips], masks[i][s:s + chunk_clips]
            if device is not None:
                x, m = x.to(device, non_blocking=True), m.to(device, non_blocking=True)
            chunks.append(predict_fn(x, m))
        logits = torch.cat(chunks)
        out_dim = logits.shape[1]
        local.append(video_mean_probs(logits, [counts[i]])[0])
    ref = next((t for t in local if t is not None), None)
    if ref is not None:
        tdev = ref.device
    elif device is not None:
        tdev = torch.device(device)
    elif distributed and dist.get_backend(group) == "nccl":
        # a rank whose shard holds no clips must still join the collectives with CUDA tensors like its peers
        tdev = torch.device("cuda", torch.cuda.current_device())
    else:
        tdev = torch.device("cpu")
    if distributed:
        od = torch.tensor([out_dim or 0], device=tdev)
        dist.all_reduce(od, op=dist.ReduceOp.MAX, group=group)
        out_dim = int(od.item())
    if not out_dim:
        return torch.full((len(videos), 0), float("nan"))
    nan_row = torch.full((out_dim,), float("nan"), device=tdev)
    rows = torch.stack([t if t is not None else nan_row for t in local]) if local else torch.empty((0, out_dim),
                                                                                                   device=tdev)
    result = torch.full((len(videos), out_dim), float("nan"), device=tdev)
    if not distributed:
        result[torch.as_tensor(mine, dtype=torch.long, device=tdev)] = rows
        return result
    shards = shard_videos(counts, world)
    width = max(len(s) for s in shards)
    padded = torch.full((width, out_dim), float("nan"), device=tdev)
    padded[:rows.shape[0]] = rows
    gathered = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(gathered, padded, group=group)  # the single collective of the path
    for r, idx in enumerate(shards):
        if idx:
            result[torch.as_tensor(idx, dtype=torch.long, device=tdev)] = gathered[r][:len(idx)]
    return result
