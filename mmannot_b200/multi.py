"""Multi-GPU side of the hot path: one process per GPU, reads sharded by read name, the annotation index
replicated, and ONE exchange step at the end -- the integer count tables and Counter's counters
(mmannot.cpp:1663) summed over the ranks.  The reference merges per-file tables in one process
(TableCount::addCounter, mmannot.cpp:1861-1876); here the same sum runs as an allreduce over NCCL.

There is no data-path collective: a read name (all its NH records, both mates) lives on exactly one rank,
so the per-read resolution never crosses GPUs.  `-y random` draws from one global rand() stream in file
order (mmannot.cpp:1711) and is therefore not shardable ("replicas only").
"""
import numpy as np

STAT_KEYS = ("n_hits", "n_reads", "n_unique", "n_ambiguous", "n_multiple", "n_unassigned", "n_rescued")


def shard_of_keys(read_key, world):
    """Rank owning each 64-bit read key (used when one file is split over the GPUs)."""
    k = np.asarray(read_key, np.uint64)
    k = (k ^ (k >> np.uint64(33))) * np.uint64(0xff51afd7ed558ccd)
    k = k ^ (k >> np.uint64(33))
    return (k % np.uint64(world)).astype(np.int64)


def read_range(rank, world, n_reads):
    """Read-name-range sharding of a name-grouped input: [first, first + count) of rank."""
    per = (n_reads + world - 1) // world
    first = min(rank * per, n_reads)
    return first, min(per, n_reads - first)


def merge_tables(res, device, group=None):
    """Sum of the per-rank results {"stats": {...}, "rows": {(mask, nh): count}} over the process group.

    1. union of the combination keys: an allgather of the (padded) local key lists  -- a few KB
    2. one allreduce(sum, int64) over the dense [union keys + 7 counters] vector
    Every rank returns the merged result."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    keys = sorted(res["rows"].keys())
    local = np.array([[m, nh] for m, nh in keys], dtype=np.uint64).reshape(-1, 2)
    n_local = torch.tensor([len(keys)], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    n_max = max(1, max(int(s[0]) for s in sizes))
    padded = np.zeros((n_max, 2), np.uint64)
    padded[:len(keys)] = local
    mine = torch.from_numpy(padded.view(np.int64)).to(device)
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    union = set()
    for r in range(world):
        g = gathered[r].cpu().numpy().view(np.uint64)[:int(sizes[r][0])]
        union.update((int(a), int(b)) for a, b in g)
    union = sorted(union)
    dense = np.zeros(len(union) + len(STAT_KEYS), np.int64)
    for i, k in enumerate(union):
        dense[i] = res["rows"].get(k, 0)
    for j, s in enumerate(STAT_KEYS):
        dense[len(union) + j] = res["stats"][s]
    t = torch.from_numpy(dense).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    out = t.cpu().numpy()
    rows = {k: int(out[i]) for i, k in enumerate(union) if int(out[i]) != 0}
    stats = {s: int(out[len(union) + j]) for j, s in enumerate(STAT_KEYS)}
    return {"stats": stats, "rows": rows}
