"""Multi-GPU plumbing of the sampling run: utterances are independent, so a batch is cut into contiguous
per-rank shards, every rank runs its own sampler loop with no communication, and ONE all-gather of the final latents
reassembles the batch (reference semantics: `NaturalSpeech2.sample` over a batch, model.py:605-696).  One process per GPU;
`torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is the only transport."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_utterances: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[begin, end) of this rank's contiguous shard; equal shards are required by the single all-gather (weak scaling:
    B per rank is fixed), so the utterance count must divide evenly."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} / world size {world_size}")
    if n_utterances % world_size:
        raise ValueError(f"{n_utterances} utterances do not split evenly over {world_size} ranks (pad the batch)")
    per = n_utterances // world_size
    return rank * per, (rank + 1) * per


def gather_latents(local: torch.Tensor, group: Optional[dist.ProcessGroup] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """All ranks' final latents `[B_local, C, T]` as one `[world * B_local, C, T]` tensor in rank order (one collective)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    local = local.contiguous()
    if out is None:
        out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)
    return out


def shard_features(world_size: int, rank: int, *tensors: torch.Tensor) -> Tuple[torch.Tensor, ...]:
    """This rank's utterances of the batch-first inputs of the device pipeline (`api.sample_from_features`: x_T [B, 100, T],
    c_padded [B, 256, T], refer_padded [B, 100, S], lengths [B], refer_lengths [B]): the condition encoders have no cross-sample op
    either (per-utterance masks, LayerNorm, attention), so the same contiguous cut applies and no collective is added."""
    if not tensors:
        return ()
    n = tensors[0].shape[0]
    if any(t.shape[0] != n for t in tensors):
        raise ValueError("all inputs must share the utterance (first) dimension")
    lo, hi = shard_bounds(n, world_size, rank)
    return tuple(t[lo:hi] for t in tensors)
