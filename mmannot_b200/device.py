"""ctypes view of the device hot path (libmmannot_b200.so, include/mmannot_b200.h).

`Annotator` plays the role of the reference's Counter for one GPU: hits go in batch by batch
(scan + addCount, mmannot.cpp:1772-1778), `finish` is the end-of-file flush plus getCounts
(mmannot.cpp:1783-1792, 1803).  There is no CPU fallback: if the CUDA library is missing or
no GPU is present this module raises.
"""
import ctypes as C
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("MMANNOT_B200_LIB") or os.path.join(_HERE, "lib", "libmmannot_b200.so")  # env: tuning builds only

STRATEGIES = {"default": 0, "unique": 1, "random": 2, "ratio": 3}
FAST_OFF = 0xFFFFFFFF  # fast_bin_shift=None: no segment answer table

EXPORTS = [
    "mma_create", "mma_destroy", "mma_last_error", "mma_load_features", "mma_alloc_pinned", "mma_free_pinned",
    "mma_submit_hits", "mma_submit_hits_device", "mma_finish_sample", "mma_reset_sample", "mma_dense_counts",
    "mma_sync", "mma_stream", "mma_timing_enable", "mma_timing_reset", "mma_timing_get", "mma_index_bytes", "mma_version", "mma_readback_bytes", "mma_dominant_kernel", "mma_index_segments", "mma_annotate_hits", "mma_annotate_intervals", "mma_pack_hits", "mma_check_packed", "mma_submit_hits_packed", "mma_device_count", "mma_warmup", "mma_export_bytes", "mma_export_table", "mma_import_tables",
    "mma_export_rows", "mma_export_head_bytes", "mma_import_tables_strided", "mma_allreduce", "mma_batch_kernel", "mma_export_table_async", "mma_restore_export",
    "mma_bam_begin", "mma_submit_bam", "mma_submit_bam_start", "mma_submit_bam_finish", "mma_bam_ref_first", "mma_bam_last_hits", "mma_bam_reserve", "mma_bam_stage",
]


class MmaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"mmannot_b200 error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    _fields_ = [("device", C.c_int32), ("strategy", C.c_int32), ("overlap", C.c_float), ("rescue_threshold", C.c_float),
                ("read_stats", C.c_int32), ("interval_stats", C.c_int32), ("n_elements", C.c_uint32),
                ("elem_line", C.c_void_p), ("elem_strand", C.c_void_p), ("elem_vicinity", C.c_void_p),
                ("n_samples", C.c_uint32), ("max_batch_hits", C.c_uint32), ("table_log2", C.c_uint32),
                ("bin_shift", C.c_uint32), ("rand_seed", C.c_uint32), ("fast_bin_shift", C.c_uint32)]


class Features(C.Structure):
    _fields_ = [("n", C.c_uint32), ("n_chr", C.c_uint32), ("chr", C.c_void_p), ("start", C.c_void_p),
                ("end", C.c_void_p), ("type", C.c_void_p), ("strand", C.c_void_p)]


class HitBatch(C.Structure):
    _fields_ = [("n", C.c_uint64), ("start", C.c_void_p), ("end", C.c_void_p), ("meta", C.c_void_p),
                ("nh", C.c_void_p), ("read_key", C.c_void_p)]


class PackedBatch(C.Structure):
    _fields_ = [("n", C.c_uint64), ("start", C.c_void_p), ("packed", C.c_void_p), ("n_runs", C.c_uint64), ("run_key", C.c_void_p),
                ("tile_run_base", C.c_void_p), ("n_escapes", C.c_uint64), ("esc_index", C.c_void_p), ("esc_end", C.c_void_p),
                ("esc_nh", C.c_void_p)]


PACK_TILE = 1024


class SampleStats(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("n_hits", "n_reads", "n_unique", "n_ambiguous", "n_multiple", "n_unassigned", "n_rescued")]


class SampleResult(C.Structure):
    _fields_ = [("stats", SampleStats), ("n_rows", C.c_uint64), ("row_mask", C.POINTER(C.c_uint64)),
                ("row_nh", C.POINTER(C.c_uint32)), ("row_count", C.POINTER(C.c_uint64))]


class Timing(C.Structure):
    _fields_ = [("ms_index", C.c_double), ("ms_batch", C.c_double), ("ms_close", C.c_double), ("ms_finish", C.c_double),
                ("launches", C.c_uint64), ("hits", C.c_uint64), ("batches", C.c_uint64), ("fast_miss", C.c_uint64),
                ("ms_bam_inflate", C.c_double), ("ms_bam_index", C.c_double), ("ms_bam_parse", C.c_double)]


_lib = None


def lib():
    """Loads the CUDA library; fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(f"{_LIB_PATH} is missing: run `make cuda` (or __graft_entry__.build()). "
                               "mmannot_b200 has no CPU fallback.")
        L = C.CDLL(_LIB_PATH)
        L.mma_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(Params)]
        L.mma_destroy.argtypes = [C.c_void_p]
        L.mma_destroy.restype = None
        L.mma_last_error.argtypes = [C.c_void_p]
        L.mma_last_error.restype = C.c_char_p
        L.mma_load_features.argtypes = [C.c_void_p, C.POINTER(Features)]
        L.mma_alloc_pinned.argtypes = [C.c_size_t]
        L.mma_alloc_pinned.restype = C.c_void_p
        L.mma_free_pinned.argtypes = [C.c_void_p]
        L.mma_free_pinned.restype = None
        L.mma_submit_hits.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(HitBatch)]
        L.mma_submit_hits_device.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(HitBatch)]
        L.mma_finish_sample.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(SampleResult)]
        L.mma_reset_sample.argtypes = [C.c_void_p, C.c_uint32]
        L.mma_dense_counts.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.mma_sync.argtypes = [C.c_void_p]
        L.mma_stream.argtypes = [C.c_void_p]
        L.mma_stream.restype = C.c_void_p
        L.mma_timing_enable.argtypes = [C.c_void_p, C.c_int]
        L.mma_timing_reset.argtypes = [C.c_void_p]
        L.mma_timing_get.argtypes = [C.c_void_p, C.POINTER(Timing)]
        L.mma_index_bytes.argtypes = [C.c_void_p]
        L.mma_index_bytes.restype = C.c_uint64
        L.mma_version.restype = C.c_char_p
        L.mma_annotate_hits.argtypes = [C.c_void_p, C.POINTER(HitBatch), C.c_void_p]
        L.mma_index_segments.argtypes = [C.c_void_p]
        L.mma_index_segments.restype = C.c_uint64
        L.mma_readback_bytes.argtypes = [C.c_void_p]
        L.mma_readback_bytes.restype = C.c_uint64
        L.mma_dominant_kernel.restype = C.c_char_p
        L.mma_alloc_pinned.argtypes = [C.c_size_t]
        L.mma_alloc_pinned.restype = C.c_void_p
        L.mma_free_pinned.argtypes = [C.c_void_p]
        L.mma_pack_hits.argtypes = [C.POINTER(HitBatch)] + [C.c_void_p] * 6 + [C.c_uint64, C.POINTER(PackedBatch)]
        L.mma_check_packed.argtypes = [C.POINTER(PackedBatch), C.c_int]
        L.mma_submit_hits_packed.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(PackedBatch)]
        L.mma_export_bytes.argtypes = [C.c_void_p]
        L.mma_export_bytes.restype = C.c_uint64
        L.mma_export_table.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.mma_import_tables.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        L.mma_export_table_async.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_uint64]
        L.mma_restore_export.argtypes = [C.c_void_p, C.c_uint32]
        L.mma_batch_kernel.argtypes = [C.c_void_p]
        L.mma_batch_kernel.restype = C.c_char_p
        L.mma_export_rows.argtypes = [C.c_void_p]
        L.mma_export_rows.restype = C.c_uint64
        L.mma_export_head_bytes.restype = C.c_uint64
        L.mma_import_tables_strided.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint64]
        L.mma_allreduce.argtypes = [C.POINTER(C.c_void_p), C.c_uint32, C.c_uint32]
        L.mma_annotate_intervals.argtypes = [C.c_void_p, C.POINTER(HitBatch), C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        _lib = L
    return _lib


def round_half_away(v):
    """C round(): what TableCount::addCounter applies to each value (mmannot.cpp:1868)."""
    return int(math.floor(v + 0.5)) if v >= 0 else -int(math.floor(-v + 0.5))


class PinnedHits:
    """Page-locked struct-of-arrays hit buffers (mma_alloc_pinned)."""

    def __init__(self, capacity):
        L = lib()
        self.capacity = int(capacity)
        self._ptrs = []
        self.arrays = {}
        for name, dt in (("start", np.uint32), ("end", np.uint32), ("meta", np.uint32), ("nh", np.uint32), ("read_key", np.uint64)):
            nbytes = self.capacity * np.dtype(dt).itemsize
            p = L.mma_alloc_pinned(nbytes)
            if not p:
                raise MemoryError("mma_alloc_pinned failed")
            self._ptrs.append(p)
            ct = C.c_uint32 if dt == np.uint32 else C.c_uint64
            self.arrays[name] = np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), shape=(self.capacity,))

    def fill(self, hits, a=0, b=None):
        b = hits.n if b is None else b
        n = b - a
        for k in ("start", "end", "meta", "nh", "read_key"):
            self.arrays[k][:n] = getattr(hits, k)[a:b]
        return n

    def batch(self, n):
        return HitBatch(n, *[self.arrays[k].ctypes.data for k in ("start", "end", "meta", "nh", "read_key")])

    def close(self):
        for p in self._ptrs:
            lib().mma_free_pinned(p)
        self._ptrs = []
        self.arrays = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PackedHits:
    """One batch in the compact transfer format (mma_packed_batch), in page-locked buffers filled by mma_pack_hits from
    wide arrays (numpy views or a PinnedHits slice).  `start` is aliased, so the wide start array must stay alive."""

    def __init__(self, start, end, meta, nh, read_key, esc_capacity=None):
        L = lib()
        n = int(len(start))
        self.n = n
        esc_capacity = max(16, n // 64) if esc_capacity is None else int(esc_capacity)
        self._ptrs = []

        def pinned(count, ct):
            p = L.mma_alloc_pinned(max(1, count) * C.sizeof(ct))
            if not p:
                raise MemoryError("mma_alloc_pinned failed")
            self._ptrs.append(p)
            return p

        self._keep = [np.ascontiguousarray(a) for a in (start, end, meta, nh, read_key)]
        wide = HitBatch(n, *[a.ctypes.data for a in self._keep])
        packed = pinned(n, C.c_uint32)
        run_key = pinned(n, C.c_uint64)
        tile = pinned((n + PACK_TILE - 1) // PACK_TILE, C.c_uint32)
        ei, ee, en = pinned(esc_capacity, C.c_uint32), pinned(esc_capacity, C.c_uint32), pinned(esc_capacity, C.c_uint32)
        self.batch = PackedBatch()
        rc = L.mma_pack_hits(C.byref(wide), packed, run_key, tile, ei, ee, en, esc_capacity, C.byref(self.batch))
        if rc != 0:
            self.close()
            raise MmaError(rc, "batch cannot be packed (too many escapes or chromosome ids): use the wide format")

    @property
    def h2d_bytes(self):
        b = self.batch
        return int(8 * b.n + 8 * b.n_runs + 4 * ((b.n + PACK_TILE - 1) // PACK_TILE) + 12 * b.n_escapes)

    def close(self):
        for p in self._ptrs:
            lib().mma_free_pinned(p)
        self._ptrs = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Annotator:
    """One device context (mma_ctx)."""

    def __init__(self, config, strategy="default", overlap=-1.0, rescue_threshold=1.0, read_stats=False,
                 interval_stats=False, n_samples=1, max_batch_hits=1 << 22, device=0, table_log2=0, bin_shift=0, rand_seed=1,
                 fast_bin_shift=0):
        self.config = config
        self._keep = (np.ascontiguousarray(config.elem_line, np.uint16), np.ascontiguousarray(config.elem_strand, np.uint8),
                      np.ascontiguousarray(config.elem_vicinity, np.uint8))
        self.strategy = STRATEGIES[strategy] if isinstance(strategy, str) else int(strategy)
        p = Params(device, self.strategy, float(overlap), float(rescue_threshold), int(bool(read_stats)), int(bool(interval_stats)),
                   len(self._keep[0]), self._keep[0].ctypes.data, self._keep[1].ctypes.data, self._keep[2].ctypes.data,
                   n_samples, int(max_batch_hits), table_log2, bin_shift, rand_seed, FAST_OFF if fast_bin_shift is None else fast_bin_shift)
        self.max_batch_hits = int(max_batch_hits)
        self._h = C.c_void_p()
        rc = lib().mma_create(C.byref(self._h), C.byref(p))
        if rc != 0:
            raise MmaError(rc, lib().mma_last_error(None).decode())

    def _check(self, rc):
        if rc != 0:
            raise MmaError(rc, lib().mma_last_error(self._h).decode())

    def load_features(self, ann):
        arrs = [np.ascontiguousarray(ann.chr, np.uint32), np.ascontiguousarray(ann.start, np.uint32),
                np.ascontiguousarray(ann.end, np.uint32), np.ascontiguousarray(ann.type, np.uint8),
                np.ascontiguousarray(ann.strand, np.uint8)]
        f = Features(len(arrs[1]), int(ann.n_chr), *[a.ctypes.data for a in arrs])
        self._check(lib().mma_load_features(self._h, C.byref(f)))

    def submit(self, sample, hits):
        """Host arrays (numpy).  Splits into max_batch_hits pieces; keeps the arrays alive until sync."""
        n = hits.n
        a = 0
        while a < n or (n == 0 and a == 0):
            b = min(n, a + self.max_batch_hits)
            part = hits.slice(a, b) if (a, b) != (0, n) else hits
            hb = HitBatch(part.n, part.start.ctypes.data, part.end.ctypes.data, part.meta.ctypes.data, part.nh.ctypes.data, part.read_key.ctypes.data)
            self._check(lib().mma_submit_hits(self._h, sample, C.byref(hb)))
            self._pending = getattr(self, "_pending", []) + [part]
            a = b
            if n == 0:
                break
        self.sync()
        self._pending = []

    def submit_batch(self, sample, hit_batch):
        """Raw mma_submit_hits on a HitBatch struct (host pointers); asynchronous."""
        self._check(lib().mma_submit_hits(self._h, sample, C.byref(hit_batch)))

    def submit_packed(self, sample, packed_batch):
        """Raw mma_submit_hits_packed on a PackedBatch struct (host pointers); asynchronous."""
        self._check(lib().mma_submit_hits_packed(self._h, sample, C.byref(packed_batch)))

    def submit_device(self, sample, hit_batch):
        """Raw mma_submit_hits_device on a HitBatch struct holding device pointers; asynchronous."""
        self._check(lib().mma_submit_hits_device(self._h, sample, C.byref(hit_batch)))

    def annotate(self, hits):
        """Per-hit element sets (IntervalList::scan alone): uint64 bitmask per hit."""
        out = np.zeros(max(hits.n, 1), np.uint64)
        hb = HitBatch(hits.n, hits.start.ctypes.data, hits.end.ctypes.data, hits.meta.ctypes.data, hits.nh.ctypes.data, hits.read_key.ctypes.data)
        self._check(lib().mma_annotate_hits(self._h, C.byref(hb), out.ctypes.data))
        return out[:hits.n]

    def sync(self):
        self._check(lib().mma_sync(self._h))

    def reset(self, sample):
        self._check(lib().mma_reset_sample(self._h, sample))

    def finish_arrays(self, sample=0, sort=True):
        """-> (stats int64[7] in SampleStats order, rows int64[n, 3] = (mask, nh, count), sorted by (mask, nh) unless
        sort=False): the result of mma_finish_sample as arrays, for callers that merge or compare tables without building
        dictionaries."""
        r = SampleResult()
        self._check(lib().mma_finish_sample(self._h, sample, C.byref(r)))
        stats = np.array([int(getattr(r.stats, k)) for k, _ in SampleStats._fields_], dtype=np.int64)
        n = int(r.n_rows)
        rows = np.empty((n, 3), np.int64)
        if n:
            rows[:, 0] = np.ctypeslib.as_array(r.row_mask, shape=(n,)).view(np.int64)
            rows[:, 1] = np.ctypeslib.as_array(r.row_nh, shape=(n,))
            rows[:, 2] = np.ctypeslib.as_array(r.row_count, shape=(n,)).view(np.int64)
            if sort:
                rows = sort_rows(rows)
        return stats, rows

    def finish(self, sample=0):
        """-> dict(stats={...}, rows={(mask, nh): count})"""
        r = SampleResult()
        self._check(lib().mma_finish_sample(self._h, sample, C.byref(r)))
        stats = {k: int(getattr(r.stats, k)) for k, _ in SampleStats._fields_}
        n = int(r.n_rows)
        if n:
            masks = np.ctypeslib.as_array(r.row_mask, shape=(n,)).tolist()
            nhs = np.ctypeslib.as_array(r.row_nh, shape=(n,)).tolist()
            counts = np.ctypeslib.as_array(r.row_count, shape=(n,)).tolist()
            rows = dict(zip(zip(masks, nhs), counts))
        else:
            rows = {}
        return {"stats": stats, "rows": rows}

    def export_bytes(self):
        return int(lib().mma_export_bytes(self._h))

    def export_table(self, sample, dev_ptr):
        """End-of-file flush, then the compacted table + counters into device memory (mma_export_bytes bytes); asynchronous."""
        self._check(lib().mma_export_table(self._h, sample, dev_ptr))

    def export_table_async(self, sample, dev_ptr, stride_bytes, rows_cap):
        self._check(lib().mma_export_table_async(self._h, sample, C.c_void_p(dev_ptr), stride_bytes, rows_cap))

    def restore_export(self, sample):
        self._check(lib().mma_restore_export(self._h, sample))

    def export_rows(self):
        return int(lib().mma_export_rows(self._h))

    def import_tables_strided(self, sample, dev_ptr, n_tables, stride_bytes, rows_cap):
        self._check(lib().mma_import_tables_strided(self._h, sample, C.c_void_p(dev_ptr), n_tables, stride_bytes, rows_cap))

    def import_tables(self, sample, dev_ptr, n_tables):
        """Replace the sample's table and counters by the sum of n_tables exported buffers laid out back to back."""
        self._check(lib().mma_import_tables(self._h, sample, dev_ptr, n_tables))

    def dense_counts(self, sample, masks, nhs, out_dev_ptr):
        m = np.ascontiguousarray(masks, np.uint64)
        h = np.ascontiguousarray(nhs, np.uint32)
        self._check(lib().mma_dense_counts(self._h, sample, m.ctypes.data, h.ctypes.data, len(m), out_dev_ptr))

    def timing_enable(self, on=True):
        self._check(lib().mma_timing_enable(self._h, int(on)))

    def timing_reset(self):
        self._check(lib().mma_timing_reset(self._h))

    def timing(self):
        t = Timing()
        self._check(lib().mma_timing_get(self._h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in Timing._fields_}

    def index_bytes(self):
        return int(lib().mma_index_bytes(self._h))

    def index_segments(self):
        return int(lib().mma_index_segments(self._h))

    def stream_ptr(self):
        """cudaStream_t (as an integer) of the context's compute stream."""
        return int(lib().mma_stream(self._h) or 0)

    def table_readback_bytes(self):
        """Bytes mma_finish_sample reads back from the device per sample."""
        return int(lib().mma_readback_bytes(self._h))

    def dominant_kernel(self):
        """The batch kernel this context launches (k_batch_lean / k_batch_fast / k_batch)."""
        return lib().mma_batch_kernel(self._h).decode() or lib().mma_dominant_kernel().decode()

    def close(self):
        if getattr(self, "_h", None):
            lib().mma_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def allreduce(annotators, sample=0):
    """mma_allreduce over the contexts of this process (one per GPU): afterwards finish() of any of them is the merged result."""
    arr = (C.c_void_p * len(annotators))(*[a._h for a in annotators])
    rc = lib().mma_allreduce(arr, len(annotators), sample)
    if rc != 0:
        raise MmaError(rc, lib().mma_last_error(annotators[0]._h).decode())


def sort_rows(rows):
    """rows int64[n, 3] = (mask, nh, count) ordered by (mask as unsigned, nh)."""
    return rows[np.lexsort((rows[:, 1], rows[:, 0].view(np.uint64)))] if len(rows) else rows


def values_by_mask(rows):
    """{(mask, nh): count} -> {mask: double}, the reference's regionCounts value (mmannot.cpp:1730)."""
    out = {}
    for (mask, nh), c in sorted(rows.items(), key=lambda kv: (kv[0][0], kv[0][1])):
        out[mask] = out.get(mask, 0.0) + (c * (1.0 / nh) if nh else float(c))
    return out


def element_vector(mask):
    return tuple(i for i in range(64) if (mask >> i) & 1)


def format_table(config, sample_names, per_sample_rows):
    """TableCount::addCounter + dump (mmannot.cpp:1861-1900): rounded values, union of rows over the
    samples, rows ordered by their element-index vectors."""
    vals = [values_by_mask(r) for r in per_sample_rows]
    masks = set()
    for v in vals:
        masks.update(v.keys())
    lines = ["Type" + "".join("\t" + s for s in sample_names)]
    for m in sorted(masks, key=element_vector):
        lines.append(config.row_name(m) + "".join("\t%d" % round_half_away(v.get(m, 0.0)) for v in vals))
    return "\n".join(lines) + "\n"
