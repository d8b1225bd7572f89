"""Compact transfer format (mma_pack_hits, host-side, no GPU needed): what the packer writes must decode back, field by
field, to the wide arrays -- the decode below restates k_expand_packed in numpy."""
import numpy as np
import pytest

from tests import fuzz
from mmannot_b200 import device

TILE = 1024


def unpack(pb, n):
    C = device.C
    as_np = lambda ptr, ct, cnt: np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(max(cnt, 1),))[:cnt].copy()
    start = as_np(pb.start, C.c_uint32, n)
    packed = as_np(pb.packed, C.c_uint32, n)
    run_key = as_np(pb.run_key, C.c_uint64, int(pb.n_runs))
    tile = as_np(pb.tile_run_base, C.c_uint32, (n + TILE - 1) // TILE)
    ne = int(pb.n_escapes)
    ei, ee, en = (as_np(p, C.c_uint32, ne) for p in (pb.esc_index, pb.esc_end, pb.esc_nh))
    length, nh = packed & 255, (packed >> 8) & 255
    end = (start + length - 1).astype(np.uint32)
    nh = nh.astype(np.uint32)
    esc = (length == 255) | (nh == 255)
    assert np.array_equal(np.nonzero(esc)[0], ei)
    end[ei] = ee
    nh[ei] = en
    chr_ = (packed >> 16) & 0x3FFF
    meta = np.where(chr_ == 0x3FFF, 0x00FFFFFF, chr_).astype(np.uint32) | (packed & np.uint32(0x80000000))
    run_start = (packed >> 30) & 1
    run_index = np.cumsum(run_start) - 1
    assert np.array_equal(tile, (np.cumsum(run_start) - run_start)[::TILE])
    return end, meta, nh, run_key[run_index]


@pytest.mark.parametrize("seed", range(4))
def test_pack_decodes_back(seed):
    rng = np.random.default_rng(8800 + seed)
    et = fuzz.make_elements(rng)
    feats = fuzz.make_features(rng, et, n_feat=200)
    hits = fuzz.make_hits(rng, feats, n_reads=5000, max_nh=(5, 300, 5, 40)[seed], messy=0.2, max_read=(60, 60, 900, 300)[seed])
    if seed == 3:
        hits.read_key[100:103] = np.uint64(0xFFFFFFFFFFFFFFFF)  # the reserved key is stored normalised
    C = device.C
    n = hits.n
    arrs = [np.ascontiguousarray(a) for a in (hits.start, hits.end, hits.meta, hits.nh, hits.read_key)]
    wide = device.HitBatch(n, *[a.ctypes.data for a in arrs])
    packed, run_key, tile = np.zeros(n, np.uint32), np.zeros(n, np.uint64), np.zeros((n + TILE - 1) // TILE, np.uint32)
    ei, ee, en = np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.uint32)
    pb = device.PackedBatch()
    rc = device.lib().mma_pack_hits(C.byref(wide), packed.ctypes.data, run_key.ctypes.data, tile.ctypes.data, ei.ctypes.data, ee.ctypes.data,
                                    en.ctypes.data, n, C.byref(pb))
    assert rc == 0
    end, meta, nh, key = unpack(pb, n)
    assert np.array_equal(end, hits.end)
    assert np.array_equal(meta, hits.meta)
    assert np.array_equal(nh, hits.nh)
    want_key = np.where(hits.read_key == np.uint64(0xFFFFFFFFFFFFFFFF), np.uint64(0xFFFFFFFFFFFFFFFE), hits.read_key)
    assert np.array_equal(key, want_key)
    runs = 1 + int(np.count_nonzero(want_key[1:] != want_key[:-1]))
    assert int(pb.n_runs) == runs
    if seed in (1, 2):
        assert int(pb.n_escapes) > 0
    # too few escape slots: refused, the caller submits the wide arrays instead
    if int(pb.n_escapes) > 1:
        rc = device.lib().mma_pack_hits(C.byref(wide), packed.ctypes.data, run_key.ctypes.data, tile.ctypes.data, ei.ctypes.data, ee.ctypes.data,
                                        en.ctypes.data, 1, C.byref(pb))
        assert rc == -4
