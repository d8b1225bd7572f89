"""Multi-GPU side of the hot path: one process per GPU, reads sharded by read name, the annotation index
replicated, and ONE exchange step at the end -- the integer count tables and Counter's counters
(mmannot.cpp:1663) summed over the ranks.  The reference merges per-file tables in one process
(TableCount::addCounter, mmannot.cpp:1861-1876); here the same sum runs as one NCCL all-gather of the
compact (key, count) rows over NVLink, every rank adding up the few thousand rows it receives.

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


def merge_on_device(ann, sample, device, group=None):
    """The merge as the GPUs do it: every rank exports its compacted table and counters into device memory
    (mma_export_table), ONE NCCL all-gather over NVLink moves them, every rank adds all of them into its own table with one
    kernel (mma_import_tables); the usual mma_finish_sample then returns the merged result.  Nothing but the control block
    crosses PCIe before that, and the host does no arithmetic.  -> (stats int64[7], rows int64[n, 3]) like finish_arrays."""
    import torch
    import torch.distributed as dist

    from .device import MmaError, lib

    world = dist.get_world_size(group)
    nbytes = ann.export_bytes()
    head = int(lib().mma_export_head_bytes())
    full_cap = (nbytes - head) // 16
    cache = getattr(ann, "_merge_buffers", None)
    if cache is None or cache[0].numel() != nbytes or cache[1].numel() != world * nbytes:
        cache = (torch.empty(nbytes, dtype=torch.uint8, device=device), torch.empty(world * nbytes, dtype=torch.uint8, device=device))
        ann._merge_buffers = cache
    mine, gathered = cache
    stream = torch.cuda.ExternalStream(ann.stream_ptr(), device=device)
    # only the live rows cross NVLink: every rank sends `cap` rows (a few thousand are in use; the table capacity is 2^16 and
    # up).  A rank holding more makes the import flag an overflow on EVERY rank (each one sees every dump's row count), and all
    # of them repeat the exchange once at full size from the dump they still hold.
    cap = min(full_cap, getattr(ann, "_merge_cap", 8192))
    with torch.cuda.stream(stream):  # the library's own compute stream: export -> all-gather -> import stay ordered on it
        optimistic = not getattr(ann, "_merge_sync", False)
        need_export = True
        while True:
            stride = (head + 16 * cap + 15) & ~15
            if need_export and optimistic:
                # nothing here waits for the GPU: flush, compaction, exchange and import queue up behind the batch kernels
                ann.export_table_async(sample, mine.data_ptr(), stride, cap)
            elif need_export:
                ann.export_table(sample, mine.data_ptr())  # (resolves the shard's deferred records first: synchronises; all rows)
            dist.all_gather_into_tensor(gathered[:world * stride], mine[:stride], group=group)
            ann.import_tables_strided(sample, gathered.data_ptr(), world, stride, cap)
            try:
                return ann.finish_arrays(sample, sort=False)
            except MmaError as e:
                if e.code == -6 and optimistic:  # some shard held deferred records: own table back, then the careful route
                    ann.restore_export(sample)
                    optimistic, need_export = False, True
                    ann._merge_sync = True
                    continue
                if e.code != -4 or cap >= full_cap:
                    raise
                # some shard holds more rows than were exchanged: once more at full size
                if optimistic:
                    ann.restore_export(sample)
                need_export = optimistic  # (the careful route already left every row in `mine`)
                cap = full_cap
                ann._merge_cap = full_cap


def merge_arrays(stats, rows, device, group=None, capacity=4096):
    """Sum of the per-rank results over the process group with ONE collective: an all-gather of the compact rows.

    stats int64[7], rows int64[n, 3] = (mask, nh, count) (Annotator.finish_arrays).  Every rank contributes a fixed-size
    vector [n, 7 counters, capacity x (mask, nh, count)] (a few tens of KB), receives everybody's, and forms the union of
    the keys and the sums itself -- the same sum TableCount::addCounter forms column by column (mmannot.cpp:1861-1876),
    without a key-agreement round before it.  If some rank holds more rows than `capacity` the exchange is repeated once
    with the capacity that fits.  Returns (stats, rows) merged, rows sorted by (mask, nh), identical on every rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    n = int(rows.shape[0])
    while True:
        payload = np.zeros(8 + 3 * capacity, np.int64)
        payload[0] = n
        payload[1:8] = stats
        if n <= capacity:
            payload[8:8 + 3 * n] = rows.reshape(-1)
        mine = torch.from_numpy(payload).to(device, non_blocking=True)
        out = torch.empty(world * payload.size, dtype=torch.int64, device=device)
        if device.type == "cuda":
            dist.all_gather_into_tensor(out, mine, group=group)
        else:  # gloo (CPU tests)
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine, group=group)
            out = torch.cat(parts)
        got = out.cpu().numpy().reshape(world, -1)
        n_max = int(got[:, 0].max())
        if n_max <= capacity:
            break
        capacity = 1 << int(n_max - 1).bit_length()
    merged_stats = got[:, 1:8].sum(axis=0)
    parts = [got[r, 8:8 + 3 * int(got[r, 0])].reshape(-1, 3) for r in range(world)]
    allrows = np.concatenate(parts) if parts else np.empty((0, 3), np.int64)
    if len(allrows):
        order = np.lexsort((allrows[:, 1], allrows[:, 0].view(np.uint64)))
        allrows = allrows[order]
        new_key = np.ones(len(allrows), bool)
        new_key[1:] = (allrows[1:, 0] != allrows[:-1, 0]) | (allrows[1:, 1] != allrows[:-1, 1])
        starts = np.nonzero(new_key)[0]
        sums = np.add.reduceat(allrows[:, 2], starts)
        merged = np.stack([allrows[starts, 0], allrows[starts, 1], sums], axis=1)
        merged = merged[merged[:, 2] != 0]
    else:
        merged = allrows
    return merged_stats, merged


def merge_tables(res, device, group=None):
    """Dictionary front-end of merge_arrays: {"stats": {...}, "rows": {(mask, nh): count}} summed over the process group;
    every rank returns the merged result."""
    keys = sorted(res["rows"].keys())
    rows = np.array([[m, nh, res["rows"][(m, nh)]] for m, nh in keys], dtype=np.uint64).reshape(-1, 3).view(np.int64)
    stats = np.array([res["stats"][s] for s in STAT_KEYS], dtype=np.int64)
    mstats, mrows = merge_arrays(stats, rows, device, group=group)
    u = mrows.view(np.uint64)
    return {"stats": {s: int(mstats[j]) for j, s in enumerate(STAT_KEYS)},
            "rows": {(int(u[i, 0]), int(u[i, 1])): int(u[i, 2]) for i in range(len(u))}}
