"""CUDA path (through the C ABI) against the CPU oracle on adversarial synthetic inputs (tests/fuzz.py):
every strategy, every -l mode, batch borders in arbitrary places, with and without the segment answer table."""
import numpy as np
import pytest

from tests import fuzz
from oracle import pyoracle

pytestmark = pytest.mark.gpu


def device_run(et, feats, hits, strategy="default", overlap=-1.0, max_batch=1 << 20, rescue_threshold=1.0, read_stats=False, **kw):
    from mmannot_b200 import device
    a = device.Annotator(et, strategy=strategy, overlap=overlap, rescue_threshold=rescue_threshold, read_stats=read_stats,
                         max_batch_hits=max_batch, **kw)
    try:
        a.load_features(feats)
        a.submit(0, hits)
        res = a.finish(0)
        res["values"] = device.values_by_mask(res["rows"])
        return res
    finally:
        a.close()


def oracle_run(et, feats, hits, **kw):
    return pyoracle.run(et.elem_line, et.elem_strand, et.elem_vicinity, feats, hits, **kw)


def check(res, ref, exact=True):
    assert set(res["values"]) == set(ref["rows"])
    for m, v in ref["rows"].items():
        if exact:
            assert res["values"][m] == v, (hex(m), res["values"][m], v)
        else:
            assert abs(res["values"][m] - v) <= 1e-9 * max(1.0, abs(v)), hex(m)
    assert res["stats"] == ref["stats"]


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("overlap", [-1.0, 0.5, 1.0, 12.0])
def test_per_hit_annotation(seed, overlap):
    """NH = 1 everywhere: the table is the histogram of the per-hit element sets."""
    rng = np.random.default_rng(1000 + seed)
    et = fuzz.make_elements(rng)
    feats = fuzz.make_features(rng, et, n_feat=int(rng.integers(50, 900)))
    hits = fuzz.make_hits(rng, feats, n_reads=20000, max_nh=1, messy=0.0)
    ref = oracle_run(et, feats, hits, overlap=overlap)
    for fast in (0, None, 3, 9):
        res = device_run(et, feats, hits, overlap=overlap, fast_bin_shift=fast, max_batch=7777)
        check(res, ref)


@pytest.mark.parametrize("seed", range(6))
def test_scan_alone(seed):
    """mma_annotate_hits (IntervalList::scan alone) against the oracle's per-hit element sets."""
    from mmannot_b200 import device
    rng = np.random.default_rng(1500 + seed)
    et = fuzz.make_elements(rng, wide=(seed == 5))
    feats = fuzz.make_features(rng, et, n_feat=int(rng.integers(50, 900)))
    hits = fuzz.make_hits(rng, feats, n_reads=30000, max_nh=1, messy=0.0)
    for overlap in (-1.0, 0.3, 0.99, 1.0, 25.0):
        ref = oracle_run(et, feats, hits, overlap=overlap, want_hit_masks=True)
        for fast in (0, None, 4):
            a = device.Annotator(et, overlap=overlap, fast_bin_shift=fast)
            try:
                a.load_features(feats)
                got = a.annotate(hits)
            finally:
                a.close()
            bad = np.nonzero(got != ref["hit_mask"])[0]
            assert len(bad) == 0, (overlap, fast, int(bad[0]), int(hits.start[bad[0]]), int(hits.end[bad[0]]), hex(int(got[bad[0]])), hex(int(ref["hit_mask"][bad[0]])))


@pytest.mark.parametrize("seed", range(4))
def test_per_hit_annotation_wide_masks(seed):
    """More than 32 Order elements: 64-bit element sets, no segment table."""
    rng = np.random.default_rng(2000 + seed)
    et = fuzz.make_elements(rng, wide=True)
    feats = fuzz.make_features(rng, et, n_feat=600)
    hits = fuzz.make_hits(rng, feats, n_reads=6000, max_nh=5)
    for overlap in (-1.0, 3.0):
        ref = oracle_run(et, feats, hits, overlap=overlap)
        check(device_run(et, feats, hits, overlap=overlap, max_batch=5000), ref)


@pytest.mark.parametrize("seed", range(8))
@pytest.mark.parametrize("batch", [1, 5, 64, 1000, 1024, 4099, 1 << 20])
def test_read_resolution_name_grouped(seed, batch):
    """default strategy: countdown over name-grouped reads incl. messy groups, batches cut anywhere."""
    rng = np.random.default_rng(3000 + seed)
    et = fuzz.make_elements(rng)
    feats = fuzz.make_features(rng, et, n_feat=300)
    n_reads = 300 if batch < 64 else 4000
    hits = fuzz.make_hits(rng, feats, n_reads=n_reads, max_nh=int(rng.choice([3, 8, 40])), messy=float(rng.choice([0.0, 0.1, 0.5])))
    overlap = float(rng.choice([-1.0, 1.0]))
    ref = oracle_run(et, feats, hits, overlap=overlap)
    check(device_run(et, feats, hits, overlap=overlap, max_batch=batch), ref)


@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("batch", [3, 700, 1 << 20])
def test_read_resolution_scattered(seed, batch):
    """Records of a read not adjacent (coordinate-sorted files): everything goes through the deferred path."""
    rng = np.random.default_rng(4000 + seed)
    et = fuzz.make_elements(rng)
    feats = fuzz.make_features(rng, et, n_feat=300)
    hits = fuzz.make_hits(rng, feats, n_reads=250 if batch < 64 else 3000, max_nh=6, messy=0.2)
    hits = fuzz.shuffle_hits(rng, hits, block=int(rng.choice([1, 16, 200])))
    ref = oracle_run(et, feats, hits)
    check(device_run(et, feats, hits, max_batch=batch), ref)


@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("strategy", ["unique", "ratio", "random"])
def test_other_strategies(seed, strategy):
    rng = np.random.default_rng(5000 + seed)
    et = fuzz.make_elements(rng)
    feats = fuzz.make_features(rng, et, n_feat=300)
    hits = fuzz.make_hits(rng, feats, n_reads=3000, max_nh=9, messy=0.2)
    if seed % 2:
        hits = fuzz.shuffle_hits(rng, hits, block=50)
    ref = oracle_run(et, feats, hits, strategy=strategy, overlap=1.0)
    for batch in (333, 1 << 20):
        check(device_run(et, feats, hits, strategy=strategy, overlap=1.0, max_batch=batch), ref, exact=(strategy != "ratio"))


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("threshold", [0.5, 0.8])
def test_rescue_threshold(seed, threshold):
    """-e with -m: rescue() needs element multiplicities, on every path (in-batch, across batches, deferred)."""
    rng = np.random.default_rng(6000 + seed)
    et = fuzz.make_elements(rng)
    feats = fuzz.make_features(rng, et, n_feat=200, extent=5000)
    hits = fuzz.make_hits(rng, feats, n_reads=2500, extent=5000, max_nh=10, messy=0.15)
    if seed % 3 == 2:
        hits = fuzz.shuffle_hits(rng, hits, block=30)
    thr = float(np.float32(threshold))
    for strategy in ("default", "ratio", "random"):
        ref = oracle_run(et, feats, hits, strategy=strategy, overlap=1.0, rescue_threshold=thr, read_stats=True)
        for batch in (97, 1 << 20):
            res = device_run(et, feats, hits, strategy=strategy, overlap=1.0, rescue_threshold=thr, read_stats=True, max_batch=batch)
            check(res, ref, exact=(strategy != "ratio"))


def test_sample_reuse_and_two_samples():
    """reset + resubmit gives the same answer; two samples on one context stay separate."""
    from mmannot_b200 import device
    rng = np.random.default_rng(7000)
    et = fuzz.make_elements(rng)
    feats = fuzz.make_features(rng, et, n_feat=300)
    h0 = fuzz.make_hits(rng, feats, n_reads=3000, max_nh=6, messy=0.2)
    h1 = fuzz.make_hits(rng, feats, n_reads=2000, max_nh=3, messy=0.0)
    r0, r1 = oracle_run(et, feats, h0), oracle_run(et, feats, h1)
    a = device.Annotator(et, n_samples=2, max_batch_hits=900)
    try:
        a.load_features(feats)
        for _ in range(2):
            a.reset(0)
            a.reset(1)
            a.submit(0, h0)
            a.submit(1, h1)
            for s, ref in ((0, r0), (1, r1)):
                res = a.finish(s)
                res["values"] = device.values_by_mask(res["rows"])
                check(res, ref)
    finally:
        a.close()


def test_empty_and_tiny_inputs():
    rng = np.random.default_rng(7100)
    et = fuzz.make_elements(rng)
    feats = fuzz.make_features(rng, et, n_feat=50)
    hits = fuzz.make_hits(rng, feats, n_reads=5, max_nh=3)
    from mmannot_b200 import host
    z = np.zeros(0, np.uint32)
    empty = host.Hits(z, z, z, z, np.zeros(0, np.uint64))
    res = device_run(et, feats, empty)
    assert res["rows"] == {} and res["stats"]["n_hits"] == 0
    check(device_run(et, feats, hits), oracle_run(et, feats, hits))
    one = hits.slice(0, 1)
    check(device_run(et, feats, one), oracle_run(et, feats, one))


@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("grid", [1, 3])
def test_multi_tile_chunks(seed, grid, monkeypatch):
    """Few blocks => every warp walks a chunk of many 128-hit tiles: runs that cross tile borders inside a chunk, runs
    that end exactly on a tile border, and the state carried from tile to tile (regression: reads ending on a tile
    border inside a chunk were dropped)."""
    monkeypatch.setenv("MMANNOT_B200_MAX_GRID", str(grid))
    rng = np.random.default_rng(9000 + seed)
    et = fuzz.make_elements(rng)
    feats = fuzz.make_features(rng, et, n_feat=500)
    hits = fuzz.make_hits(rng, feats, n_reads=40000, max_nh=(70 if seed & 1 else 6), messy=(0.0 if seed < 2 else 0.1))
    for strategy in ("default", "unique", "ratio"):
        ref = oracle_run(et, feats, hits, strategy=strategy)
        for batch in (1 << 20, 33333):
            res = device_run(et, feats, hits, strategy=strategy, max_batch=batch)
            check(res, ref, exact=(strategy != "ratio"))
    ref = oracle_run(et, feats, hits, rescue_threshold=0.6, read_stats=True)
    check(device_run(et, feats, hits, rescue_threshold=0.6, read_stats=True, max_batch=50000), ref)


def test_natural_grid_two_tiles_per_warp():
    """Enough hits that the full-size grid gives each warp more than one tile (> 148 SMs x 5 blocks x 8 warps x 128 hits)."""
    rng = np.random.default_rng(4242)
    et = fuzz.make_elements(rng)
    feats = fuzz.make_features(rng, et, n_feat=800)
    hits = fuzz.make_hits(rng, feats, n_reads=450000, max_nh=5, messy=0.02)
    assert hits.n > 148 * 5 * 8 * 128
    ref = oracle_run(et, feats, hits)
    check(device_run(et, feats, hits, max_batch=1 << 21), ref)


@pytest.mark.parametrize("seed", range(5))
def test_packed_transfer_format(seed):
    """mma_pack_hits + mma_submit_hits_packed (8 B/hit + 8 B/run over PCIe, expanded on the device) must give exactly what
    the wide arrays give: long reads and NH >= 255 (escapes), unknown chromosomes, empty-CIGAR intervals, cut batches."""
    from mmannot_b200 import device
    rng = np.random.default_rng(7000 + seed)
    et = fuzz.make_elements(rng)
    feats = fuzz.make_features(rng, et, n_feat=400)
    hits = fuzz.make_hits(rng, feats, n_reads=30000, max_nh=(6 if seed < 3 else 300), messy=0.1, max_read=(60 if seed != 1 else 700))
    strategy = ("default", "default", "ratio", "default", "unique")[seed]
    ref = oracle_run(et, feats, hits, strategy=strategy)
    batch = (1 << 20, 25000, 4099, 100000, 1 << 20)[seed]
    a = device.Annotator(et, strategy=strategy, max_batch_hits=batch)
    keep = []
    try:
        a.load_features(feats)
        n_esc = 0
        for lo in range(0, hits.n, batch):
            part = hits.slice(lo, min(hits.n, lo + batch))
            ph = device.PackedHits(part.start, part.end, part.meta, part.nh, part.read_key, esc_capacity=part.n)
            n_esc += int(ph.batch.n_escapes)
            assert ph.h2d_bytes < 24 * part.n
            keep.append(ph)
            a.submit_packed(0, ph.batch)
        res = a.finish(0)
    finally:
        a.close()
        for ph in keep:
            ph.close()
    res["values"] = device.values_by_mask(res["rows"])
    check(res, ref, exact=(strategy != "ratio"))
    if seed in (1, 3):
        assert n_esc > 0


def test_packed_refuses_what_does_not_fit():
    from mmannot_b200 import device
    n = 1000
    start = np.arange(1, n + 1, dtype=np.uint32) * 1000
    end = start + 5000  # every hit needs an escape
    meta = np.zeros(n, np.uint32); nh = np.ones(n, np.uint32); key = np.arange(n, dtype=np.uint64)
    with pytest.raises(device.MmaError):
        device.PackedHits(start, end, meta, nh, key, esc_capacity=10)


def test_export_import_tables_on_device():
    """The device-side merge used across GPUs, on one GPU: a sample's exported table imported twice must give every count
    and counter doubled; imported once, the table itself."""
    import torch
    from mmannot_b200 import device
    rng = np.random.default_rng(31337)
    et = fuzz.make_elements(rng)
    feats = fuzz.make_features(rng, et, n_feat=300)
    hits = fuzz.make_hits(rng, feats, n_reads=20000, max_nh=5, messy=0.1)
    ref = oracle_run(et, feats, hits)
    a = device.Annotator(et, max_batch_hits=1 << 20)
    try:
        a.load_features(feats)
        a.submit(0, hits)
        nb = a.export_bytes()
        buf = torch.empty(2 * nb, dtype=torch.uint8, device="cuda")
        a.export_table(0, buf.data_ptr())
        a.sync()
        buf[nb:] = buf[:nb]
        torch.cuda.synchronize()
        a.import_tables(0, buf.data_ptr(), 1)
        once = a.finish(0)
        a.import_tables(0, buf.data_ptr(), 2)
        twice = a.finish(0)
    finally:
        a.close()
    once["values"] = device.values_by_mask(once["rows"])
    check(once, ref)
    assert {k: 2 * v for k, v in once["rows"].items()} == twice["rows"]
    assert {k: 2 * v for k, v in once["stats"].items()} == twice["stats"]


@pytest.mark.parametrize("groups", ["0", "1"])
@pytest.mark.parametrize("seed", range(3))
def test_runs_of_k_times_nh_records(seed, groups, monkeypatch):
    """Paired-end shaped input: every name is a run of 2 x NH (sometimes 3 x NH) records carrying NH, i.e. two (three) reads
    back to back (mmannot.cpp:1673-1698).  Both variants of k_batch_fast (runs resolved group by group in parallel /
    serial walker), natural and capped grids, cut batches, plus some irregular names."""
    monkeypatch.setenv("MMANNOT_B200_GROUPS", groups)
    if seed == 2:
        monkeypatch.setenv("MMANNOT_B200_MAX_GRID", "2")
    rng = np.random.default_rng(5100 + seed)
    et = fuzz.make_elements(rng)
    feats = fuzz.make_features(rng, et, n_feat=400)
    base = fuzz.make_hits(rng, feats, n_reads=30000, max_nh=(4, 9, 40)[seed], messy=(0.0, 0.03, 0.0)[seed])
    # every name k times in a row
    heads = np.nonzero(np.concatenate([[True], base.read_key[1:] != base.read_key[:-1]]))[0]
    ends = np.concatenate([heads[1:], [base.n]])
    idx = []
    for a, b in zip(heads, ends):
        k = 3 if (a % 11 == 0) else 2
        for _ in range(k):
            idx.extend(range(a, b))
    idx = np.array(idx)
    from mmannot_b200 import host
    hits = host.Hits(base.start[idx], base.end[idx], base.meta[idx], base.nh[idx], base.read_key[idx])
    ref = oracle_run(et, feats, hits)
    for batch in (1 << 21, 20011):
        check(device_run(et, feats, hits, max_batch=batch), ref)


@pytest.mark.parametrize("defer", ["0", "1"])
@pytest.mark.parametrize("seed", range(4))
def test_defer_mode(seed, defer, monkeypatch):
    """k_batch_lean's DEFER mode (every record with NH > 1 goes to the deferred list; one sort by read key; the records of a
    name taken in file order by selection) pinned on and off, on name-grouped, messy and shuffled input, with and without
    rescue(), batches cut anywhere, natural and capped grids: the table must not depend on the mode."""
    monkeypatch.setenv("MMANNOT_B200_DEFER", defer)
    if seed == 3:
        monkeypatch.setenv("MMANNOT_B200_MAX_GRID", "2")
    rng = np.random.default_rng(6100 + seed)
    et = fuzz.make_elements(rng)
    feats = fuzz.make_features(rng, et, n_feat=400)
    hits = fuzz.make_hits(rng, feats, n_reads=25000, max_nh=(3, 12, 60, 8)[seed], messy=(0.0, 0.2, 0.05, 0.3)[seed])
    if seed in (1, 2):
        hits = fuzz.shuffle_hits(rng, hits, block=int((1, 64)[seed - 1]))
    ref = oracle_run(et, feats, hits)
    for batch in (1 << 21, 10007, 129):
        check(device_run(et, feats, hits, max_batch=batch), ref)
    ref = oracle_run(et, feats, hits, rescue_threshold=0.5, read_stats=True)
    check(device_run(et, feats, hits, rescue_threshold=0.5, read_stats=True, max_batch=30011), ref)


def test_defer_mode_is_found_without_help():
    """Coordinate-sorted input with a single batch per sample: the first sample's counters switch the context to DEFER mode for
    the next one (and name-grouped input switches it back); results are those of the oracle every time."""
    from mmannot_b200 import device
    rng = np.random.default_rng(6200)
    et = fuzz.make_elements(rng)
    feats = fuzz.make_features(rng, et, n_feat=400)
    grouped = fuzz.make_hits(rng, feats, n_reads=30000, max_nh=6, messy=0.0)
    order = np.lexsort((grouped.start, grouped.meta & 0xFFFFFF))
    from mmannot_b200 import host
    sorted_hits = host.Hits(*[np.ascontiguousarray(getattr(grouped, k)[order]) for k in ("start", "end", "meta", "nh", "read_key")])
    a = device.Annotator(et, max_batch_hits=1 << 21)
    try:
        a.load_features(feats)
        for hits in (sorted_hits, sorted_hits, grouped, grouped, sorted_hits):
            ref = oracle_run(et, feats, hits)
            a.reset(0)
            a.submit(0, hits)
            res = a.finish(0)
            res["values"] = device.values_by_mask(res["rows"])
            check(res, ref)
    finally:
        a.close()
