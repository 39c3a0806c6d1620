"""Multi-GPU host logic on CPU: world_size-2 gloo processes shard a variable-length batch, run the
path's module API on their shard (the oracle stands in for the CUDA path) and gather on the host.
The result must equal the unsharded run (utterances are independent: SURVEY.md §8e)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

import styletts_zs_b200 as stz


def test_shard_utterances_balanced_and_complete():
    lens = [5, 300, 17, 512, 64, 64, 16, 211, 90]
    for world in (1, 2, 4, 8):
        sh = stz.shard_utterances(lens, world)
        assert sorted(i for s in sh for i in s) == list(range(len(lens)))
        assert max(len(s) for s in sh) - min(len(s) for s in sh) <= 1
        for s in sh:
            assert [lens[i] for i in s] == sorted((lens[i] for i in s), reverse=True)
    assert stz.shard_utterances([], 2) == [[], []]


def _oracle_compute(cfg, w):
    from oracle.model import OraclePath
    o = OraclePath(cfg, w)

    def compute(text, mask, prompt, pmask, noise, seed=None, first_utterance=0):
        z = o.sample_style(text, prompt, 2, 2.0, text_mask=mask, prompt_mask=pmask, noise=noise, seed=seed,
                           first_utterance=first_utterance)
        return z, o.predict_duration(text, z, text_mask=mask)
    return compute


def _seeded(inp, seed=99):
    out = {k: v for k, v in inp.items() if k != "noise"}
    out["seed"] = seed
    return out


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    cfg = stz.TINY
    w = stz.init_weights(cfg, 0)
    inp = stz.synthetic_inputs(cfg, 5, 14, steps=2, seed=11, var_len=(3, 14))
    compute = _oracle_compute(cfg, w)
    res = stz.synthesize_sharded(compute, inp, rank, world)
    res_seeded = stz.synthesize_sharded(compute, _seeded(inp), rank, world)   # on-"device" noise from global indices
    # the product form: outputs land in a shared host mapping (no collective on the data path), every rank re-orders its own rows
    shards = stz.shard_utterances(inp["lens"].tolist(), world)
    shard_T = [max(int(inp["lens"][i]) for i in sh) if sh else 1 for sh in shards]
    out = stz.SharedHostOutputs(f"test{port}", 5, 14, cfg.n_style, cfg.d_style, rank, world, dist.barrier)

    def compute_out(text, mask, prompt, pmask, noise, out_style=None, out_dur=None):
        return compute(text, mask, prompt, pmask, noise)      # no out= support: the sharder copies into the slab
    res_shm = stz.synthesize_sharded_shm(compute_out, stz.take_shard(inp, shards[rank]) if shards[rank] else None, shards,
                                         shard_T, out, dist.barrier)
    if rank == 0:
        # the ordered result every rank scattered into == rank 0 re-ordering the slabs by itself (the older form)
        a_style, a_dur = out.assemble(shards, shard_T)
        assert torch.equal(a_style, res_shm[0]) and torch.equal(a_dur, res_shm[1])
        q.put((res[0], res[1], res_seeded[0], res_seeded[1], res_shm[0].clone(), res_shm[1].clone()))
    else:
        assert res is None and res_seeded is None and res_shm is None
    out.close(dist.barrier)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world_size_2_matches_single_process():
    cfg = stz.TINY
    w = stz.init_weights(cfg, 0)
    inp = stz.synthetic_inputs(cfg, 5, 14, steps=2, seed=11, var_len=(3, 14))
    ref_style, ref_dur = stz.synthesize_sharded(_oracle_compute(cfg, w), inp, 0, 1)
    full = _oracle_compute(cfg, w)(inp["text_emb"], inp["text_mask"], inp["prompt_feats"], inp["prompt_mask"], inp["noise"])
    assert torch.allclose(ref_style, full[0], atol=2e-5) and torch.equal(ref_dur, full[1].to(torch.int32))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    style, dur, style_seeded, dur_seeded, style_shm, dur_shm = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert torch.allclose(style, ref_style, atol=2e-5)
    assert torch.equal(dur, ref_dur)
    assert torch.equal(style_shm, style) and torch.equal(dur_shm, dur)       # shared-memory gather == object gather
    assert not os.path.exists(f"/dev/shm/stz_test{port}_style")               # rank 0 unlinked the mapping
    # seed mode: the two ranks drew their utterances' noise from global indices == the unsharded seeded run
    full_seeded = _oracle_compute(cfg, w)(inp["text_emb"], inp["text_mask"], inp["prompt_feats"], inp["prompt_mask"], None, seed=99)
    assert torch.allclose(style_seeded, full_seeded[0], atol=2e-5)
    assert torch.equal(dur_seeded, full_seeded[1].to(torch.int32))


def test_seeded_shards_draw_the_unsharded_noise():
    """inputs["seed"] instead of a noise tensor: each shard draws its utterances' counter-based noise from their GLOBAL
    indices, so the gathered result equals the unsharded seeded run (no noise tensor is sliced or moved)."""
    from oracle.model import OraclePath
    cfg = stz.TINY
    o = OraclePath(cfg, stz.init_weights(cfg, 0))

    def compute(text, mask, prompt, pmask, noise, seed=None, first_utterance=0):
        z = o.sample_style(text, prompt, 2, 2.0, text_mask=mask, prompt_mask=pmask, noise=noise, seed=seed,
                           first_utterance=first_utterance)
        return z, o.predict_duration(text, z, text_mask=mask)
    inp = stz.synthetic_inputs(cfg, 5, 14, steps=2, seed=11, var_len=(3, 14))
    inp = {k: v for k, v in inp.items() if k != "noise"}
    inp["seed"] = 99
    full = compute(inp["text_emb"], inp["text_mask"], inp["prompt_feats"], inp["prompt_mask"], None, seed=99)
    shards = stz.shard_utterances(inp["text_mask"].sum(1).tolist(), 2)
    style = torch.zeros_like(full[0])
    for r in range(2):                                   # both ranks' local halves, no process group needed
        sh = stz.take_shard(inp, shards[r])
        z, _ = compute(sh["text_emb"], sh["text_mask"], sh["prompt_feats"], sh.get("prompt_mask"), None, seed=99,
                       first_utterance=shards[r])
        style[torch.tensor(shards[r])] = z
    assert torch.allclose(style, full[0], atol=2e-5)
    one = stz.synthesize_sharded(compute, inp, 0, 1)     # the sharder's own seed path (world 1)
    assert torch.allclose(one[0], full[0], atol=2e-5) and torch.equal(one[1], full[1].to(torch.int32))
