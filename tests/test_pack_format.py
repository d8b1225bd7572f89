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
    assert device.lib().mma_check_packed(C.byref(pb), 0) == 0  # what mma_submit_hits_packed checks on every call
    assert device.lib().mma_check_packed(C.byref(pb), 1) == 0  # ... and the O(n) check of the run-start bits and the escapes
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


def _pack(hits):
    C = device.C
    n = hits.n
    arrs = [np.ascontiguousarray(a) for a in (hits.start, hits.end, hits.meta, hits.nh, hits.read_key)]
    wide = device.HitBatch(n, *[a.ctypes.data for a in arrs])
    bufs = dict(packed=np.zeros(n, np.uint32), run_key=np.zeros(n, np.uint64), tile=np.zeros((n + TILE - 1) // TILE, np.uint32),
                ei=np.zeros(n, np.uint32), ee=np.zeros(n, np.uint32), en=np.zeros(n, np.uint32))
    pb = device.PackedBatch()
    rc = device.lib().mma_pack_hits(C.byref(wide), *[bufs[k].ctypes.data for k in ("packed", "run_key", "tile", "ei", "ee", "en")], n, C.byref(pb))
    assert rc == 0
    return pb, bufs, arrs


def test_check_packed_refuses_inconsistent_batches():
    """mma_check_packed (host side, no GPU): every way the struct can disagree with itself is refused -- by the O(tiles + escapes)
    check that mma_submit_hits_packed runs, or, for the run-start bits and missing escapes, by the deep check."""
    rng = np.random.default_rng(99)
    et = fuzz.make_elements(rng)
    feats = fuzz.make_features(rng, et, n_feat=200)
    hits = fuzz.make_hits(rng, feats, n_reads=4000, max_nh=300, messy=0.2, max_read=400)
    C = device.C
    check = lambda pb, deep: device.lib().mma_check_packed(C.byref(pb), deep)
    pb, bufs, _keep = _pack(hits)
    n, n_tiles, n_esc = int(pb.n), len(bufs["tile"]), int(pb.n_escapes)
    assert n_tiles >= 3 and n_esc >= 2 and check(pb, 0) == 0 and check(pb, 1) == 0

    def broken(key, index, value, deep_only=False):
        old = bufs[key][index]
        bufs[key][index] = value
        shallow, deep = check(pb, 0), check(pb, 1)
        bufs[key][index] = old
        assert deep == -1, (key, index)
        assert shallow == (0 if deep_only else -1), (key, index)

    broken("tile", 0, 1)                                     # does not start at 0
    broken("tile", 2, int(bufs["tile"][1]) - 1)              # falls
    broken("tile", 1, int(bufs["tile"][0]) + TILE + 1)       # more runs than the tile has hits
    broken("tile", n_tiles - 1, int(pb.n_runs) + 1)          # beyond n_runs
    broken("tile", 1, int(bufs["tile"][1]) + 1, deep_only=(int(bufs["tile"][1]) + 1 <= int(bufs["tile"][2])))  # off by one: only the bits tell
    broken("ei", 1, int(bufs["ei"][0]))                      # not strictly increasing
    broken("ei", n_esc - 1, n)                               # beyond the batch
    first_run_start = int(np.nonzero((bufs["packed"][1:] >> 30) & 1)[0][0]) + 1
    broken("packed", first_run_start, int(bufs["packed"][first_run_start]) & ~(1 << 30), deep_only=True)  # a run start goes missing
    broken("packed", 0, int(bufs["packed"][0]) & ~(1 << 30))                                              # record 0 must start a run
    not_escaped = int(np.nonzero(((bufs["packed"] & 255) != 255) & (((bufs["packed"] >> 8) & 255) != 255))[0][5])
    broken("packed", not_escaped, int(bufs["packed"][not_escaped]) | 255, deep_only=True)                 # escaped but not listed
    # counts
    for field, value in (("n_runs", 0), ("n_runs", n + 1), ("n_escapes", n + 1)):
        old = getattr(pb, field)
        setattr(pb, field, value)
        assert check(pb, 0) == -1, field
        setattr(pb, field, old)
    old = pb.run_key
    pb.run_key = None
    assert check(pb, 0) == -1
    pb.run_key = old
    pb.n = 0
    assert check(pb, 0) == 0 and check(pb, 1) == 0  # an empty batch is fine
