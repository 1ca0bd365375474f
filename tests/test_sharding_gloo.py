"""CPU, world_size 2 over gloo: the N>1 host path of the database scan -- residue-balanced
sharding, gathering per-subject results back into caller order, and the top-k merge.  The scan
itself needs a GPU; here every rank's shard is scored by the CPU oracle (test infrastructure)
so that the distributed plumbing is checked end to end without a device."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import psb_data

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import psb_data as pd
        from oracle import oracle as orc
        from parasail_rs_b200 import sharding
        query = pd.random_seq(2001, 0, 60)
        cat, off = pd.protein_db(2002, 2003, 400, query=query, planted_frac=0.05)
        shards, shard_of = sharding.make_shards(cat, off, world)
        my_cat, my_off, ids = shards[rank]
        omat = orc.Matrix.from_table(pd.BLOSUM62_ALPHABET, pd.blosum62_table())
        loc = orc.align_batch(query, np.array([0, len(query)]), my_cat, my_off, omat, mode=orc.SW, open=10, gap=1,
                              shared_query=True)
        local = {k: loc[k] for k in ("score", "end_query", "end_ref")}
        full = sharding.gather_results(local, ids, len(off) - 1, dist, dst=0)
        top_ids, top_sc = sharding.merge_topk(ids, loc["score"], 7, dist)
        # step time: max over ranks, as bench.py reports it
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            ref = orc.align_batch(query, np.array([0, len(query)]), cat, off, omat, mode=orc.SW, open=10, gap=1,
                                  shared_query=True)
            ok = all(np.array_equal(full[k], ref[k]) for k in local)
            order = np.lexsort((np.arange(len(ref["score"])), -ref["score"].astype(np.int64)))[:7]
            ok = ok and np.array_equal(top_ids, order) and np.array_equal(top_sc, ref["score"][order])
            ok = ok and float(t.item()) == float(world)
            loads = np.bincount(shard_of, weights=np.diff(off), minlength=world)
            ok = ok and (loads.max() - loads.min()) <= np.diff(off).max()
            ret.put(bool(ok))
    finally:
        dist.destroy_process_group()


def test_two_rank_scan_plumbing():
    import __graft_entry__ as g
    g.build()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(timeout=240) for p in procs]
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert ret.get(timeout=5) is True


def test_local_shard_partitions_everything():
    import __graft_entry__ as g
    g.build()
    from parasail_rs_b200 import sharding
    cat, off = psb_data.protein_db(11, 12, 3000)
    for world in (1, 2, 3, 8):
        shards, shard_of = sharding.make_shards(cat, off, world)
        seen = np.concatenate([s[2] for s in shards])
        assert np.array_equal(np.sort(seen), np.arange(len(off) - 1))
        for c, o, ids in shards:
            for t in (0, len(ids) // 2, len(ids) - 1):
                assert np.array_equal(c[o[t]: o[t + 1]], cat[off[ids[t]]: off[ids[t] + 1]])
