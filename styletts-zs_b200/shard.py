"""Multi-GPU partitioning of the hot path: utterances are independent, so a batch
is split across ranks with NO collective on the data path; results are gathered on the host.

The reference has no parallel code to cite (/root/reference/README.md:11-16); BASELINE.json's
north_star fixes the scheme: "Utterance batches shard across the 8 GPUs of one box with no
collective on the hot path, only a host-side gather".

One process per GPU (torchrun); `group` is a host-side (gloo) process group used only for the final
gather of CPU tensors.  `compute` is any callable with the module API's semantics — the product
passes StyleTTSZSPath.synthesize_host, the CPU tests pass the oracle.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch


def shard_utterances(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Length-sorted round-robin: rank r gets utterances order[r::world] (longest first), so every
    rank sees a similar mix of lengths and its padded batch is as tight as its longest utterance."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    return [order[r::world_size] for r in range(world_size)]


def take_shard(inputs: Dict[str, torch.Tensor], idx: List[int]) -> Dict[str, torch.Tensor]:
    """Rows `idx` of a batch dict (text_emb, text_mask, prompt_feats, prompt_mask, noise), with the
    text axis trimmed to the shard's longest utterance."""
    ii = torch.tensor(idx, dtype=torch.long)
    mask = inputs["text_mask"][ii]
    t_max = max(int(mask.sum(1).max()), 1) if len(idx) else 1
    out = {"text_emb": inputs["text_emb"][ii][:, :t_max].contiguous(), "text_mask": mask[:, :t_max].contiguous(),
           "prompt_feats": inputs["prompt_feats"][ii].contiguous()}
    if inputs.get("noise") is not None:
        out["noise"] = inputs["noise"][:, ii].contiguous()
    if inputs.get("prompt_mask") is not None:
        out["prompt_mask"] = inputs["prompt_mask"][ii].contiguous()
    return out


def synthesize_sharded(compute: Callable[..., Tuple[torch.Tensor, torch.Tensor]], inputs: Dict[str, torch.Tensor],
                       rank: int, world_size: int, group=None) -> Optional[Tuple[torch.Tensor, torch.Tensor]]:
    """Runs `compute(text_emb, text_mask, prompt_feats, prompt_mask, noise) -> (style [n,K,Ds], dur [n,t])`
    on this rank's shard and gathers on the host.  Rank 0 returns (style [B,K,Ds], dur [B,T]) in the
    original utterance order (durations zero-padded back to T); other ranks return None.

    Noise: either `inputs["noise"]` ([slices,B,K,Ds], sliced per shard) or `inputs["seed"]` (int) — then
    `compute(..., None, seed=seed, first_utterance=idx)` is called with the shard's GLOBAL utterance indices and
    draws the counter-based noise itself (include/stz.h: stz_set_noise_utterances), so every utterance gets the
    noise it would get in the unsharded batch without any noise tensor crossing the host."""
    B, T = inputs["text_mask"].shape
    lengths = inputs["text_mask"].sum(1).tolist()
    shards = shard_utterances(lengths, world_size)
    idx = shards[rank]
    local = None
    if idx:
        sh = take_shard(inputs, idx)
        if inputs.get("noise") is not None:
            style, dur = compute(sh["text_emb"], sh["text_mask"], sh["prompt_feats"], sh.get("prompt_mask"), sh["noise"])
        else:
            style, dur = compute(sh["text_emb"], sh["text_mask"], sh["prompt_feats"], sh.get("prompt_mask"), None,
                                 seed=int(inputs["seed"]), first_utterance=list(idx))
        local = (style.cpu(), dur.cpu())
    if world_size == 1:
        gathered = [local]
    else:
        import torch.distributed as dist
        gathered = [None] * world_size if rank == 0 else None
        dist.gather_object(local, gathered, dst=0, group=group)
    if rank != 0:
        return None
    style_out, dur_out = None, torch.zeros(B, T, dtype=torch.int32)
    for r, res in enumerate(gathered):
        if res is None:
            continue
        st, du = res
        if style_out is None:
            style_out = torch.zeros(B, *st.shape[1:], dtype=st.dtype)
        ii = torch.tensor(shards[r], dtype=torch.long)
        style_out[ii] = st
        dur_out[ii, :du.shape[1]] = du.to(torch.int32)
    return style_out, dur_out
