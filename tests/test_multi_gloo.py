"""Multi-rank side of the path on CPU: world_size-2 gloo process group, read-name sharding and the table merge
(mmannot_b200/multi.py).  The per-rank results come from the CPU oracle here (the checker standing in for a GPU rank);
the merged table must equal the oracle on the unsharded input."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import common, fuzz
from oracle import pyoracle
from mmannot_b200 import host, multi


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case(seed):
    rng = np.random.default_rng(seed)
    et = fuzz.make_elements(rng)
    feats = fuzz.make_features(rng, et, n_feat=300)
    hits = fuzz.make_hits(rng, feats, n_reads=3000, max_nh=6, messy=0.2)
    return et, feats, hits


def _oracle_rows(et, feats, hits, strategy):
    r = pyoracle.run(et.elem_line, et.elem_strand, et.elem_vicinity, feats, hits, strategy=strategy, overlap=1.0)
    # integer rows like the device returns them: {(mask, nh): count}; default/unique have nh = 0
    return {"stats": r["stats"], "rows": {(m, 0): int(round(v)) for m, v in r["rows"].items()}}


def _worker(rank, world, port, seed, strategy, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, common.ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    et, feats, hits = _case(seed)
    owner = multi.shard_of_keys(hits.read_key, world)
    sel = np.nonzero(owner == rank)[0]
    mine = host.Hits(hits.start[sel], hits.end[sel], hits.meta[sel], hits.nh[sel], hits.read_key[sel])
    local = _oracle_rows(et, feats, mine, strategy)
    merged = multi.merge_tables(local, torch.device("cpu"))
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), np.array(sorted((m, nh, c) for (m, nh), c in merged["rows"].items()), dtype=np.uint64).reshape(-1, 3))
    np.save(os.path.join(out_dir, "stats%d.npy" % rank), np.array([merged["stats"][k] for k in multi.STAT_KEYS], dtype=np.int64))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("strategy", ["default", "unique"])
def test_sharded_merge_equals_unsharded(tmp_path, strategy):
    world, seed = 2, 424242
    mp.spawn(_worker, args=(world, _free_port(), seed, strategy, str(tmp_path)), nprocs=world, join=True)
    et, feats, hits = _case(seed)
    ref = _oracle_rows(et, feats, hits, strategy)
    want = np.array(sorted((m, nh, c) for (m, nh), c in ref["rows"].items()), dtype=np.uint64).reshape(-1, 3)
    for rank in range(world):
        got = np.load(os.path.join(tmp_path, "rank%d.npy" % rank))
        assert np.array_equal(got, want)
        stats = np.load(os.path.join(tmp_path, "stats%d.npy" % rank))
        assert stats.tolist() == [ref["stats"][k] for k in multi.STAT_KEYS]


def test_read_range_covers_everything_once():
    for n in (0, 1, 7, 100, 12345):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                first, count = multi.read_range(r, world, n)
                seen.extend(range(first, first + count))
            assert seen == list(range(n))


def test_key_sharding_keeps_names_together():
    rng = np.random.default_rng(7)
    keys = rng.integers(0, 2**63, 1000, dtype=np.uint64)
    keys = np.repeat(keys, 3)
    for world in (2, 4, 8):
        owner = multi.shard_of_keys(keys, world)
        assert owner.min() >= 0 and owner.max() < world
        assert np.array_equal(owner[0::3], owner[1::3]) and np.array_equal(owner[0::3], owner[2::3])
        assert len(np.unique(owner)) == world
