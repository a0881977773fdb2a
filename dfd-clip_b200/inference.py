"""Video-level scoring driver: the caller loop of the reference's ``inference.py:107-156`` around the hot path.

* ``predict_from_host``: clips in pinned host memory -> logits in host memory, H2D copies double-buffered on a copy
  stream so they overlap the encoder of the previous chunk (the reference does a blocking ``.to(device)`` /
  ``.to("cpu")`` per chunk, inference.py:116-118).
* ``shard_videos`` / ``score_videos``: videos are independent units, sharded across ranks by clip count; each rank
  scores its videos (softmax per clip, mean over the clips of a video, inference.py:121,140) and ONE all_gather of
  the per-video scores assembles the result on every rank (the reference gathers once per video, :147-149).
  No collective runs inside the clip loop.
"""
import torch


class HostClipPipeline:
    """Reusable double-buffered host->device->host pipeline for ``Detector.predict``."""

    def __init__(self, detector, chunk_clips=16):
        self.det = detector
        self.chunk = int(chunk_clips)
        self.dev = next(detector.decoder.parameters()).device
        if self.dev.type != "cuda":
            raise RuntimeError("HostClipPipeline needs the detector on a CUDA device")
        self.copy_stream = torch.cuda.Stream(self.dev)
        self._bufs = None

    def _buffers(self, x_host, m_host):
        shape = (self.chunk,) + tuple(x_host.shape[1:])
        if self._bufs is None or self._bufs[0][0].shape != shape:
            self._bufs = [(torch.empty(shape, dtype=torch.float32, device=self.dev),
                           torch.empty((self.chunk, m_host.shape[1]), dtype=torch.bool, device=self.dev))
                          for _ in range(2)]
        return self._bufs

    def __call__(self, x_host, m_host, out_host=None):
        """x_host fp32 [B,T,3,R,R], m_host bool [B,T] (pinned for real overlap). Returns logits of task 0 on the host
        (fp32 [B, out_dim], pinned); the call returns after the last D2H copy has completed."""
        n = x_host.shape[0]
        det, dev, main = self.det, self.dev, torch.cuda.current_stream(self.dev)
        out_dim = det.out_dim[0]
        if out_host is None:
            out_host = torch.empty((n, out_dim), dtype=torch.float32).pin_memory()
        if n == 0:
            return out_host
        bufs = self._buffers(x_host, m_host)
        starts = list(range(0, n, self.chunk))
        copied = [torch.cuda.Event() for _ in starts]
        consumed = [torch.cuda.Event() for _ in starts]
        self.copy_stream.wait_stream(main)
        for i, s in enumerate(starts):
            e = min(s + self.chunk, n)
            xb, mb = bufs[i % 2]
            with torch.cuda.stream(self.copy_stream):
                if i >= 2:
                    self.copy_stream.wait_event(consumed[i - 2])  # buffer free again
                xb[:e - s].copy_(x_host[s:e], non_blocking=True)
                mb[:e - s].copy_(m_host[s:e], non_blocking=True)
                copied[i].record(self.copy_stream)
            main.wait_event(copied[i])
            logits, _ = det.predict(xb[:e - s], mb[:e - s])
            consumed[i].record(main)
            out_host[s:e].copy_(logits[0], non_blocking=True)
        main.synchronize()
        return out_host


_PIPELINES = {}


def predict_from_host(detector, x_host, m_host, chunk_clips=16):
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
