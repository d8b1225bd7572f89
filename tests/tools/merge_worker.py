"""Worker of tests/test_gpu_multi.py (launched under torch.distributed.run, one rank per GPU): every rank annotates its
shard (read names dealt out by hash) and the ranks merge on the devices (multi.merge_on_device: export / NCCL all-gather
of the live rows / import); the merged table must be the oracle's on the UNSHARDED input.  Exit code 0 = all good."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from mmannot_b200 import device, multi, host
    from oracle import pyoracle
    from tests import fuzz

    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    failures = []
    for seed, strategy, overlap, shuffle in ((1, "default", -1.0, False), (2, "unique", 1.0, False), (3, "ratio", -1.0, False), (4, "default", -1.0, True)):
        rng = np.random.default_rng(9000 + seed)  # same stream on every rank
        et = fuzz.make_elements(rng)
        feats = fuzz.make_features(rng, et, n_feat=400)
        hits = fuzz.make_hits(rng, feats, n_reads=20000, max_nh=8, messy=0.2)
        if shuffle:
            hits = fuzz.shuffle_hits(rng, hits, block=50)
        ref = pyoracle.run(et.elem_line, et.elem_strand, et.elem_vicinity, feats, hits, strategy=strategy, overlap=overlap)
        mine = multi.shard_of_keys(hits.read_key, world) == rank
        part = host.Hits(*[np.ascontiguousarray(getattr(hits, k)[mine]) for k in ("start", "end", "meta", "nh", "read_key")])
        for cap in (8192, 16):  # (16: fewer rows than the shards hold -> the exchange is repeated at full size)
            ann = device.Annotator(et, strategy=strategy, overlap=overlap, max_batch_hits=7001, device=local)
            ann._merge_cap = cap
            try:
                ann.load_features(feats)
                ann.submit(0, part)
                stats, rows = multi.merge_on_device(ann, 0, dev)
            finally:
                ann.close()
            u = device.sort_rows(rows).view(np.uint64)
            got = {}
            for m, nh, c in u:
                got[int(m)] = got.get(int(m), 0.0) + float(c) * (1.0 / int(nh) if nh else 1.0)
            st = {k: int(v) for k, v in zip(multi.STAT_KEYS, stats)}
            ok = set(got) == set(ref["rows"]) and all(abs(got[m] - v) <= 1e-9 * max(1.0, abs(v)) for m, v in ref["rows"].items()) and st == ref["stats"]
            if not ok:
                failures.append((seed, strategy, cap, st, ref["stats"]))
    dist.barrier()
    dist.destroy_process_group()
    if failures:
        print("rank %d FAILED: %s" % (rank, failures), file=sys.stderr)
        return 1
    print("rank %d ok" % rank)
    return 0


if __name__ == "__main__":
    sys.exit(main())
