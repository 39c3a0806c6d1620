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

import math
import os
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


class SharedHostOutputs:
    """The host-side gather without a gather: one /dev/shm-backed mapping per job, shared by the ranks of one box.  Rank r's
    device->host copies land directly in slab r (the mapping is registered as pinned memory when CUDA is present); every
    rank then places ITS OWN rows at their positions in the caller's utterance order inside the same mapping
    (``scatter_own``: the re-ordering is host work of B / world_size rows per rank, in parallel), and after a barrier rank 0
    holds the complete, ordered result (``ordered``) — no collective, no pickling, no serial re-ordering pass.
    (``assemble`` is the older form: rank 0 re-orders every slab by itself.)

    ``tag`` names the job and must be the same on every rank (bench.py uses MASTER_PORT); ``barrier`` is a callable that
    synchronises the ranks (``dist.barrier`` of a gloo / nccl group; a no-op for world_size 1)."""

    def __init__(self, tag: str, B: int, T: int, K: int, Ds: int, rank: int, world_size: int, barrier: Callable[[], None],
                 pin: bool = True):
        self.rank, self.world, self.B, self.T, self.K, self.Ds = rank, world_size, B, T, K, Ds
        self.n_max = math.ceil(B / world_size)
        self._paths = [f"/dev/shm/stz_{tag}_style", f"/dev/shm/stz_{tag}_dur"]
        n_slab_style, n_slab_dur = world_size * self.n_max * K * Ds, world_size * self.n_max * T
        n_style, n_dur = n_slab_style + B * K * Ds, n_slab_dur + B * T          # slabs, then the ordered result
        if rank == 0:
            for path, nbytes in zip(self._paths, (4 * n_style, 4 * n_dur)):
                with open(path, "wb") as f:
                    f.truncate(nbytes)
        barrier()
        self._map_style = torch.from_file(self._paths[0], shared=True, size=n_style, dtype=torch.float32)
        self._map_dur = torch.from_file(self._paths[1], shared=True, size=n_dur, dtype=torch.int32)
        self.style = self._map_style[:n_slab_style].view(world_size, self.n_max, K, Ds)
        self.dur = self._map_dur[:n_slab_dur].view(world_size, self.n_max * T)
        self.style_all = self._map_style[n_slab_style:].view(B, K, Ds)          # caller's utterance order
        self.dur_all = self._map_dur[n_slab_dur:].view(B, T)
        self._registered = []
        if pin and torch.cuda.is_available():
            rt = torch.cuda.cudart()
            for t in (self.style, self.dur):     # only the slabs are DMA targets
                if int(rt.cudaHostRegister(t.data_ptr(), t.numel() * t.element_size(), 0)) == 0:
                    self._registered.append(t.data_ptr())
        self.pinned = len(self._registered) == 2
        barrier()

    def slab(self, n: int, t: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """This rank's output buffers: style [n,K,Ds] fp32 and durations [n,t] int32 (contiguous)."""
        return self.style[self.rank, :n], self.dur[self.rank, :n * t].view(n, t)

    def scatter_own(self, shards: List[List[int]], shard_T: List[int]) -> None:
        """This rank's rows -> their positions in the caller's utterance order (durations zero-padded to T)."""
        idx = shards[self.rank]
        if not idx:
            return
        ii = torch.tensor(idx, dtype=torch.long)
        n, t = len(idx), shard_T[self.rank]
        self.style_all.index_copy_(0, ii, self.style[self.rank, :n])
        rows = torch.zeros(n, self.T, dtype=torch.int32)
        rows[:, :t] = self.dur[self.rank, :n * t].view(n, t)
        self.dur_all.index_copy_(0, ii, rows)

    def ordered(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """After every rank's ``scatter_own`` and a barrier: (style [B,K,Ds], dur [B,T]) in the caller's order — views of the
        shared mapping (valid until ``close`` or the next batch; ``.clone()`` to keep them)."""
        return self.style_all, self.dur_all

    def assemble(self, shards: List[List[int]], shard_T: List[int]) -> Tuple[torch.Tensor, torch.Tensor]:
        """Rank 0, after the barrier: results in the caller's utterance order (durations zero-padded to T)."""
        style = torch.empty(self.B, self.K, self.Ds)
        dur = torch.zeros(self.B, self.T, dtype=torch.int32)
        for r, idx in enumerate(shards):
            if not idx:
                continue
            ii = torch.tensor(idx, dtype=torch.long)
            n, t = len(idx), shard_T[r]
            style.index_copy_(0, ii, self.style[r, :n])
            dur[ii, :t] = self.dur[r, :n * t].view(n, t)
        return style, dur

    def close(self, barrier: Optional[Callable[[], None]] = None):
        if self._registered:
            rt = torch.cuda.cudart()
            for ptr in self._registered:
                rt.cudaHostUnregister(ptr)
            self._registered = []
        self.style = self.dur = self.style_all = self.dur_all = self._map_style = self._map_dur = None
        if barrier is not None:
            barrier()
        if self.rank == 0:
            for path in self._paths:
                try:
                    os.unlink(path)
                except OSError:
                    pass


def synthesize_sharded_shm(compute: Callable[..., Tuple[torch.Tensor, torch.Tensor]], shard_inputs: Optional[Dict[str, torch.Tensor]],
                           shards: List[List[int]], shard_T: List[int], out: SharedHostOutputs,
                           barrier: Callable[[], None]) -> Optional[Tuple[torch.Tensor, torch.Tensor]]:
    """The product form of ``synthesize_sharded``: this rank's shard (``shard_inputs`` = ``take_shard(inputs, shards[rank])``,
    prepared by the caller: in a server every rank receives its own utterances) runs through
    ``compute(text_emb, text_mask, prompt_feats, prompt_mask, noise, out_style=..., out_dur=...)`` with the output buffers
    inside the shared host mapping; every rank scatters its own rows into the ordered result; a barrier.  Returns
    (style [B,K,Ds], dur [B,T]) on rank 0 — views of the shared mapping, see ``SharedHostOutputs.ordered`` — and None elsewhere."""
    idx = shards[out.rank]
    if idx:
        o_style, o_dur = out.slab(len(idx), shard_T[out.rank])
        sh = shard_inputs
        res = compute(sh["text_emb"], sh["text_mask"], sh["prompt_feats"], sh.get("prompt_mask"), sh.get("noise"),
                      out_style=o_style, out_dur=o_dur)
        if res is not None and res[0] is not None and res[0].data_ptr() != o_style.data_ptr():   # a compute without out= support
            o_style.copy_(res[0])
            o_dur.copy_(res[1])
    out.scatter_own(shards, shard_T)
    barrier()
    if out.rank != 0:
        return None
    return out.ordered()
