"""Multi-GPU plumbing for the database scan: one process per GPU, the database dealt to
ranks by residue count, no collective on the data path.  Only per-subject results (or each
rank's top-k) travel, over torch.distributed (NCCL on the GPU box, gloo in the CPU tests).
"""
import numpy as np

from . import shard_plan


def local_shard(cat, off, shard_of, rank):
    """(residues, offsets, global subject ids) of the subjects assigned to `rank`."""
    off = np.asarray(off, dtype=np.int64)
    ids = np.nonzero(np.asarray(shard_of) == rank)[0]
    lens = (off[ids + 1] - off[ids]).astype(np.int64)
    my_off = np.zeros(len(ids) + 1, dtype=np.int64)
    my_off[1:] = np.cumsum(lens)
    my_cat = np.empty(int(my_off[-1]), dtype=np.uint8)
    # copy maximal runs of consecutive ids at once
    if len(ids):
        breaks = np.nonzero(np.diff(ids) != 1)[0] + 1
        starts = np.concatenate([[0], breaks])
        ends = np.concatenate([breaks, [len(ids)]])
        for a, b in zip(starts, ends):
            my_cat[my_off[a]: my_off[b]] = cat[off[ids[a]]: off[ids[b - 1] + 1]]
    return my_cat, my_off, ids


def make_shards(cat, off, world):
    """plan + cut: list of (residues, offsets, global ids), one per rank"""
    shard_of = shard_plan(off, world)
    return [local_shard(cat, off, shard_of, r) for r in range(world)], shard_of


def gather_results(local, global_ids, n_total, dist=None, dst=0):
    """Assemble per-subject int32 arrays (dict name -> array over the rank's subjects) into
    caller-order arrays on rank `dst` (returns None elsewhere).  Without a process group the
    local arrays are scattered directly."""
    import torch
    names = sorted(local)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        out = {k: np.zeros(n_total, dtype=np.int32) for k in names}
        for k in names:
            out[k][global_ids] = local[k]
        return out
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([len(global_ids)], dtype=torch.int64, device=dev))
    counts = [int(c.item()) for c in counts]
    width = max(counts)
    payload = torch.zeros((len(names) + 1, width), dtype=torch.int64, device=dev)
    payload[0, : len(global_ids)] = torch.from_numpy(np.asarray(global_ids, dtype=np.int64)).to(dev)
    for t, k in enumerate(names):
        payload[t + 1, : len(global_ids)] = torch.from_numpy(np.asarray(local[k], dtype=np.int64)).to(dev)
    bufs = [torch.zeros_like(payload) for _ in range(world)] if rank == dst else None
    dist.gather(payload, bufs, dst=dst)
    if rank != dst:
        return None
    out = {k: np.zeros(n_total, dtype=np.int32) for k in names}
    for r in range(world):
        b = bufs[r].cpu().numpy()
        ids = b[0, : counts[r]]
        for t, k in enumerate(names):
            out[k][ids] = b[t + 1, : counts[r]].astype(np.int32)
    return out


def merge_topk(local_ids, local_scores, k, dist=None):
    """Global top-k from each rank's candidates (global subject ids + scores): every rank
    contributes its own k best, one all_gather of k x 2 integers, merge on the host.  Ties go to
    the smaller subject id (same rule as psb_batch_topk)."""
    import torch
    local_ids = np.asarray(local_ids, dtype=np.int64)
    local_scores = np.asarray(local_scores, dtype=np.int64)
    order = np.lexsort((local_ids, -local_scores))[:k]
    cand = np.full((2, k), -1, dtype=np.int64)
    cand[0, : len(order)] = local_ids[order]
    cand[1, : len(order)] = local_scores[order]
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        mine = torch.from_numpy(cand).to(dev)
        bufs = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
        dist.all_gather(bufs, mine)
        cand = np.concatenate([b.cpu().numpy() for b in bufs], axis=1)
    keep = cand[0] >= 0
    ids, sc = cand[0][keep], cand[1][keep]
    order = np.lexsort((ids, -sc))[:k]
    return ids[order], sc[order].astype(np.int32)
