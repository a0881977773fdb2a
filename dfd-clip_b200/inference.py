"""Video-level scoring driver: the caller loop of the reference's ``inference.py:107-156`` around the hot path.

* ``predict_from_host``: clips in pinned host memory -> logits in host memory; the encoder runs chunk by chunk with
  the H2D copies double-buffered on a copy stream, the decoder once per batch (the reference does a blocking
  ``.to(device)`` / ``.to("cpu")`` around every chunk, inference.py:116-118).
* ``shard_videos`` / ``score_videos``: videos are independent units, sharded across ranks by clip count; each rank
  scores its videos (softmax per clip, mean over the clips of a video, inference.py:121,140) and ONE all_gather of
  the per-video scores assembles the result on every rank (the reference gathers once per video, :147-149).
  No collective runs inside the clip loop.
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
    """softmax per clip, then mean over each video's clips (inference.py:121,140). clip_logits [sum(counts), O]."""
    probs = clip_logits.float().softmax(dim=-1)
    counts = torch.as_tensor(list(clip_counts), device=probs.device, dtype=torch.long)
    video_id = torch.repeat_interleave(torch.arange(len(clip_counts), device=probs.device), counts)
    sums = torch.zeros((len(clip_counts), probs.shape[1]), dtype=probs.dtype, device=probs.device)
    sums.index_add_(0, video_id, probs)
    return sums / counts.clamp_min(1).unsqueeze(1).to(probs.dtype)


def score_videos_batched(detector, videos, masks, batch_clips=64, group=None):
    """The same result as ``score_videos`` for host-resident videos, computed the way the hardware likes it: the
    clips of ALL of this rank's videos are packed into batches of ``batch_clips`` (video boundaries do not matter:
    clips are independent) and pushed through ``HostClipPipeline`` (H2D double-buffered under the encoder, decoder once
    per batch, one D2H of the logits); the per-video mean of clip probabilities is a segment mean at the end and ONE
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
        x_host = torch.cat([videos[i] for i in mine])
        m_host = torch.cat([masks[i] for i in mine])
        if not x_host.is_pinned():
            x_host, m_host = x_host.pin_memory(), m_host.pin_memory()
        pipe = HostClipPipeline(detector)
        pipe.MAX_BATCH = int(batch_clips)
        logits = pipe(x_host, m_host).to(dev)
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
    tdev = ref.device if ref is not None else (torch.device(device) if device is not None else torch.device("cpu"))
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
