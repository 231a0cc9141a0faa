"""Multi-GPU sharding of the video batch: one process per GPU, no collective on the hot path.

Videos are independent, so the batch is split contiguously over ranks, each rank
decodes its shard locally, and a single all-gather of the int32 token matrix (+ lengths) at the end
assembles the captions -- latency-bound (about 1 MB per rank at 8192 videos x 31 tokens).
The reference has no distributed code; this is the only collective in the package.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of ``n`` videos for ``rank``; the first ``n % world`` ranks get one extra."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_captions_equal(tokens: torch.Tensor, lengths: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Fast path when every rank holds the same [n, L] shape (the benchmark's weak-scaling shards): a single
    fixed-shape ``all_gather_into_tensor`` with no host synchronisation."""
    if not (dist.is_available() and dist.is_initialized()):
        return tokens, lengths
    world = dist.get_world_size(group)
    n, L = tokens.shape
    buf = torch.empty(n, L + 1, dtype=torch.int32, device=tokens.device)
    buf[:, :L] = tokens
    buf[:, L] = lengths
    out = torch.empty(world * n, L + 1, dtype=torch.int32, device=tokens.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    return out[:, :L], out[:, L]


def gather_captions(tokens: torch.Tensor, lengths: torch.Tensor, pad_id: int, group=None
                    ) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather ragged per-rank results: tokens [n_r, L_r] int, lengths [n_r] -> ([N, L], [N]) on every rank.

    Rows are right-padded with ``pad_id`` (the reference pads with START, video_captioning_model.py:288-300)
    to the global maximum length, and shards are padded to the largest shard so one equal-size
    all_gather suffices (NCCL needs equal sizes).  Works on CUDA (nccl) and CPU (gloo) tensors.
    """
    if not (dist.is_available() and dist.is_initialized()):
        return tokens, lengths
    world = dist.get_world_size(group)
    dev = tokens.device
    meta = torch.tensor([tokens.shape[0], tokens.shape[1]], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    counts = [int(m[0]) for m in metas]
    L = max(int(m[1]) for m in metas)
    nmax = max(counts)
    buf = torch.full((nmax, L + 1), pad_id, dtype=torch.int32, device=dev)
    buf[: tokens.shape[0], : tokens.shape[1]] = tokens.to(torch.int32)
    buf[: tokens.shape[0], L] = lengths.to(torch.int32)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    toks = torch.cat([o[:c, :L] for o, c in zip(out, counts)], dim=0).to(torch.int64)
    lens = torch.cat([o[:c, L] for o, c in zip(out, counts)], dim=0).to(torch.int64)
    return toks, lens


class ShardedCaptioner:
    """Data-parallel caption generation: ``generate`` takes this rank's shard (or the full batch with
    ``already_sharded=False``) and returns the gathered result on every rank.

    Beam rows are exactly the rows of an unsharded call (row i == the B=1 call on video i).  Greedy: the reference's
    batched loop stops at the first step where EVERY row of the batch emits END (decoder.py:275), so a shard may stop
    at an earlier column than the whole batch would; gathered rows are right-padded with END, which decodes to the same
    caption (``decode_caption`` drops END) but the columns after a row's first END are not those of an unsharded call.
    A rank whose shard is empty (fewer videos than ranks) contributes zero rows and still takes part in the gather."""

    def __init__(self, model, group=None):
        self.model = model
        self.group = group

    @property
    def rank(self) -> int:
        return dist.get_rank(self.group) if dist.is_available() and dist.is_initialized() else 0

    @property
    def world_size(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def generate(self, video_features: torch.Tensor, start_token_id: int, end_token_id: int, max_length: int = 20,
                 video_mask: Optional[torch.Tensor] = None, method: str = "greedy", already_sharded: bool = True,
                 gather: bool = True, **kwargs) -> Dict[str, torch.Tensor]:
        if not already_sharded:
            lo, hi = shard_bounds(video_features.shape[0], self.world_size, self.rank)
            video_features = video_features[lo:hi]
            video_mask = None if video_mask is None else video_mask[lo:hi]
        # (an empty shard -- fewer videos than ranks -- yields empty results and still takes part in the gather)
        out = self.model.generate(video_features, start_token_id, end_token_id, max_length=max_length,
                                  video_mask=video_mask, method=method, **kwargs)
        toks = out["generated_tokens"]
        if "lengths" in out:
            lens = out["lengths"]
        else:
            lens = torch.full((toks.shape[0],), toks.shape[1], dtype=torch.int64, device=toks.device)
        if gather and self.world_size > 1:
            pad = start_token_id if method == "beam" else end_token_id
            toks, lens = gather_captions(toks, lens, pad, self.group)
        return {"generated_tokens": toks, "lengths": lens}
