"""ctypes view of the host front-end (libmmannot_host.so, include/mmannot_b200_host.h).

Config / annotation / alignment decoding -- the producers of the packed buffers that
cross the device boundary.  Mirrors Config (mmannot.cpp:219-471), the IntervalList
constructor (mmannot.cpp:1094-1290) and the SAM/BAM readers (mmannot.cpp:1339-1650).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "lib", "libmmannot_host.so")

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(f"{_LIB_PATH} is missing: run `make host` (or __graft_entry__.build())")
        L = C.CDLL(_LIB_PATH)
        L.mmh_last_error.restype = C.c_char_p
        L.mmh_config_load.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.mmh_config_free.argtypes = [C.c_void_p]
        L.mmh_config_n_elements.argtypes = [C.c_void_p]
        L.mmh_config_n_elements.restype = C.c_uint32
        L.mmh_config_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mmh_config_name.argtypes = [C.c_void_p, C.c_uint32, C.c_char_p, C.c_size_t]
        L.mmh_config_name.restype = C.c_size_t
        L.mmh_config_order_echo.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        L.mmh_config_order_echo.restype = C.c_size_t
        L.mmh_annotation_build.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_void_p)]
        L.mmh_annotation_free.argtypes = [C.c_void_p]
        for name, res in [("n", C.c_uint32), ("n_chr", C.c_uint32), ("n_genes", C.c_uint64), ("n_lines", C.c_uint64),
                          ("chr", C.c_void_p), ("start", C.c_void_p), ("end", C.c_void_p), ("type", C.c_void_p),
                          ("strand", C.c_void_p), ("warnings", C.c_char_p)]:
            f = getattr(L, "mmh_annotation_" + name)
            f.argtypes = [C.c_void_p]
            f.restype = res
        L.mmh_annotation_id.argtypes = [C.c_void_p, C.c_uint32]
        L.mmh_annotation_id.restype = C.c_char_p
        L.mmh_annotation_chr_name.argtypes = [C.c_void_p, C.c_uint32]
        L.mmh_annotation_chr_name.restype = C.c_char_p
        L.mmh_reader_open.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_char, C.POINTER(C.c_void_p)]
        L.mmh_reader_close.argtypes = [C.c_void_p]
        L.mmh_reader_next.argtypes = [C.c_void_p, C.c_uint64] + [C.c_void_p] * 5
        L.mmh_reader_next.restype = C.c_uint64
        L.mmh_reader_records.argtypes = [C.c_void_p]
        L.mmh_reader_records.restype = C.c_uint64
        L.mmh_reader_warnings.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        L.mmh_reader_warnings.restype = C.c_size_t
        L.mmh_reader_key_collision.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        L.mmh_reader_key_collision.restype = C.c_size_t
        L.mmh_name_key.argtypes = [C.c_char_p, C.c_size_t]
        L.mmh_name_key.restype = C.c_uint64
        _lib = L
    return _lib


class HostError(RuntimeError):
    pass


def _err():
    return HostError(lib().mmh_last_error().decode("utf-8", "replace"))


class Config:
    """Parsed configuration file; element i = i-th item of the flattened Order section."""

    def __init__(self, path):
        self._h = C.c_void_p()
        if lib().mmh_config_load(os.fsencode(path), C.byref(self._h)) != 0:
            raise _err()
        self.n_elements = lib().mmh_config_n_elements(self._h)
        self.elem_line = np.zeros(self.n_elements, np.uint16)
        self.elem_strand = np.zeros(self.n_elements, np.uint8)
        self.elem_vicinity = np.zeros(self.n_elements, np.uint8)
        lib().mmh_config_tables(self._h, self.elem_line.ctypes.data, self.elem_strand.ctypes.data, self.elem_vicinity.ctypes.data)
        buf = C.create_string_buffer(4096)
        self.names = []
        for i in range(self.n_elements):
            lib().mmh_config_name(self._h, i, buf, 4096)
            self.names.append(buf.value.decode())

    def order_echo(self):
        buf = C.create_string_buffer(1 << 16)
        lib().mmh_config_order_echo(self._h, buf, 1 << 16)
        return buf.value.decode()

    def row_name(self, mask):
        """Table row label of an element set (TableCount::dump, mmannot.cpp:1889-1893)."""
        return "--".join(self.names[i] for i in range(self.n_elements) if (int(mask) >> i) & 1)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().mmh_config_free(self._h)
            self._h = None


def _view(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype)
    ct = {np.uint32: C.c_uint32, np.uint8: C.c_uint8}[dtype]
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,)).copy()


class Annotation:
    """Typed intervals in reference order (the feature buffer of mma_load_features)."""

    def __init__(self, config, gtf_path, upstream=1000, downstream=1000):
        self.config = config
        self._h = C.c_void_p()
        if lib().mmh_annotation_build(config._h, os.fsencode(gtf_path), upstream, downstream, C.byref(self._h)) != 0:
            raise _err()
        L = lib()
        self.n = L.mmh_annotation_n(self._h)
        self.n_chr = L.mmh_annotation_n_chr(self._h)
        self.n_genes = L.mmh_annotation_n_genes(self._h)
        self.n_lines = L.mmh_annotation_n_lines(self._h)
        self.chr = _view(L.mmh_annotation_chr(self._h), self.n, np.uint32)
        self.start = _view(L.mmh_annotation_start(self._h), self.n, np.uint32)
        self.end = _view(L.mmh_annotation_end(self._h), self.n, np.uint32)
        self.type = _view(L.mmh_annotation_type(self._h), self.n, np.uint8)
        self.strand = _view(L.mmh_annotation_strand(self._h), self.n, np.uint8)
        self.warnings = L.mmh_annotation_warnings(self._h).decode("utf-8", "replace")

    def ids(self):
        return [lib().mmh_annotation_id(self._h, i).decode() for i in range(self.n)]

    def chromosomes(self):
        return [lib().mmh_annotation_chr_name(self._h, i).decode() for i in range(self.n_chr)]

    def __del__(self):
        if getattr(self, "_h", None):
            lib().mmh_annotation_free(self._h)
            self._h = None


class Hits:
    """Struct-of-arrays hit batch (mma_hit_batch)."""

    __slots__ = ("start", "end", "meta", "nh", "read_key")

    def __init__(self, start, end, meta, nh, read_key):
        self.start = np.ascontiguousarray(start, np.uint32)
        self.end = np.ascontiguousarray(end, np.uint32)
        self.meta = np.ascontiguousarray(meta, np.uint32)
        self.nh = np.ascontiguousarray(nh, np.uint32)
        self.read_key = np.ascontiguousarray(read_key, np.uint64)

    @property
    def n(self):
        return int(self.start.shape[0])

    def slice(self, a, b):
        return Hits(self.start[a:b], self.end[a:b], self.meta[a:b], self.nh[a:b], self.read_key[a:b])

    @staticmethod
    def concat(parts):
        return Hits(*[np.concatenate([getattr(p, k) for p in parts]) for k in Hits.__slots__])


def read_hits(annotation, path, strandedness="F", fmt=0, batch=1 << 20, collision=None):
    """Decode a whole SAM/BAM file into one Hits object (plus the decoder's warnings).  `collision`: a list that receives the
    decoder's read-key verification result ("" or the first two different names sharing a key)."""
    L = lib()
    h = C.c_void_p()
    if L.mmh_reader_open(annotation._h, os.fsencode(path), fmt, strandedness.encode()[0:1], C.byref(h)) != 0:
        raise _err()
    parts = []
    try:
        while True:
            arrs = [np.empty(batch, np.uint32) for _ in range(4)] + [np.empty(batch, np.uint64)]
            n = L.mmh_reader_next(h, batch, *[a.ctypes.data for a in arrs])
            if n == 0:
                break
            parts.append(Hits(*[a[:n] for a in arrs]))
        buf = C.create_string_buffer(1 << 20)
        L.mmh_reader_warnings(h, buf, 1 << 20)
        warnings = buf.value.decode("utf-8", "replace")
        if collision is not None:
            L.mmh_reader_key_collision(h, buf, 1 << 20)
            collision.append(buf.value.decode("utf-8", "replace"))
    finally:
        L.mmh_reader_close(h)
    if not parts:
        z = np.zeros(0, np.uint32)
        return Hits(z, z, z, z, np.zeros(0, np.uint64)), warnings
    return Hits.concat(parts), warnings


def name_key(name):
    b = name.encode() if isinstance(name, str) else name
    return int(lib().mmh_name_key(b, len(b)))


class ElementTable:
    """Element table without a config file behind it (e.g. loaded from a fixture)."""

    def __init__(self, elem_line, elem_strand, elem_vicinity, names=None):
        self.elem_line = np.ascontiguousarray(elem_line, np.uint16)
        self.elem_strand = np.ascontiguousarray(elem_strand, np.uint8)
        self.elem_vicinity = np.ascontiguousarray(elem_vicinity, np.uint8)
        self.n_elements = len(self.elem_line)
        self.names = list(names) if names is not None else ["e%d" % i for i in range(self.n_elements)]

    def row_name(self, mask):
        return "--".join(self.names[i] for i in range(self.n_elements) if (int(mask) >> i) & 1)


class FeatureArrays:
    """Feature buffer without a GTF behind it."""

    def __init__(self, chr, start, end, type, strand, n_chr):
        self.chr = np.ascontiguousarray(chr, np.uint32)
        self.start = np.ascontiguousarray(start, np.uint32)
        self.end = np.ascontiguousarray(end, np.uint32)
        self.type = np.ascontiguousarray(type, np.uint8)
        self.strand = np.ascontiguousarray(strand, np.uint8)
        self.n_chr = int(n_chr)
        self.n = len(self.start)


class _SynthReads(C.Structure):
    _fields_ = [("max_nh", C.c_uint32), ("paired", C.c_int32), ("flip_mate2", C.c_int32), ("rna_seq", C.c_int32),
                ("p_in_feature", C.c_double), ("p_same_class", C.c_double)]


class Synth:
    """Synthetic annotation + reads of a benchmark shape ("tair10", "hs38", "flybase6")."""

    def __init__(self, shape, seed, gene_scale=1.0, max_nh=20, paired=False, flip_mate2=False, rna_seq=False,
                 p_in_feature=0.5, p_same_class=0.6):
        L = lib()
        L.mmh_synth_create.argtypes = [C.c_char_p, C.c_uint64, C.c_double, C.POINTER(C.c_void_p)]
        L.mmh_synth_free.argtypes = [C.c_void_p]
        L.mmh_synth_n_genes.argtypes = [C.c_void_p]
        L.mmh_synth_n_genes.restype = C.c_uint64
        L.mmh_synth_write_annotation.argtypes = [C.c_void_p, C.c_char_p]
        L.mmh_synth_write_bam.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_uint64, C.POINTER(_SynthReads), C.c_int]
        L.mmh_synth_count_hits.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(_SynthReads)]
        L.mmh_synth_count_hits.restype = C.c_uint64
        L.mmh_synth_fill_hits.argtypes = [C.c_void_p, C.c_void_p, C.c_char, C.c_uint64, C.c_uint64, C.POINTER(_SynthReads), C.c_uint64] + [C.c_void_p] * 5
        L.mmh_synth_fill_hits.restype = C.c_uint64
        self.shape = shape
        self.spec = _SynthReads(max_nh, int(paired), int(flip_mate2), int(rna_seq), p_in_feature, p_same_class)
        self._h = C.c_void_p()
        if L.mmh_synth_create(shape.encode(), seed, gene_scale, C.byref(self._h)) != 0:
            raise _err()
        self.n_genes = L.mmh_synth_n_genes(self._h)

    def write_annotation(self, path):
        lib().mmh_synth_write_annotation(self._h, os.fsencode(path))

    def write_bam(self, path, first_read, n_reads, coordinate_sorted=False, headerless=False, straddle=False):
        if lib().mmh_synth_write_bam(self._h, os.fsencode(path), first_read, n_reads, C.byref(self.spec), int(coordinate_sorted) | (2 if headerless else 0) | (4 if straddle else 0)) != 0:
            raise _err()

    def write_bam_parallel(self, path, first_read, n_reads, threads):
        """One name-grouped BAM written as `threads` parts side by side (BGZF files concatenate: only the first part has the header)."""
        from concurrent.futures import ThreadPoolExecutor
        import shutil
        threads = max(1, min(threads, n_reads // 100000 or 1))
        step = (n_reads + threads - 1) // threads
        parts = [(first_read + t * step, min(step, n_reads - t * step)) for t in range(threads) if t * step < n_reads]
        names = [path if i == 0 else "%s.part%d" % (path, i) for i in range(len(parts))]
        with ThreadPoolExecutor(len(parts)) as ex:
            list(ex.map(lambda i: self.write_bam(names[i], parts[i][0], parts[i][1], headerless=(i > 0)), range(len(parts))))
        with open(path, "ab") as out:
            for nm in names[1:]:
                with open(nm, "rb") as f:
                    shutil.copyfileobj(f, out, 16 << 20)
                os.unlink(nm)

    def count_hits(self, first_read, n_reads):
        return int(lib().mmh_synth_count_hits(self._h, first_read, n_reads, C.byref(self.spec)))

    def fill_hits(self, annotation, strandedness, first_read, n_reads, out=None, offset=0):
        """Packed hits of the read range.  `out` = dict of preallocated arrays (filled at `offset`) or None."""
        if out is None:
            n = self.count_hits(first_read, n_reads)
            out = {"start": np.empty(n, np.uint32), "end": np.empty(n, np.uint32), "meta": np.empty(n, np.uint32),
                   "nh": np.empty(n, np.uint32), "read_key": np.empty(n, np.uint64)}
        cap = len(out["start"]) - offset
        ptrs = [out[k][offset:].ctypes.data for k in ("start", "end", "meta", "nh", "read_key")]
        n = lib().mmh_synth_fill_hits(self._h, annotation._h, strandedness.encode()[0:1], first_read, n_reads, C.byref(self.spec), cap, *ptrs)
        return out, int(n)

    def hits(self, annotation, strandedness, first_read, n_reads, threads=1):
        """All hits of the range as one Hits object; `threads` > 1 fills disjoint slices in parallel."""
        if threads <= 1 or n_reads < 4 * threads:
            out, n = self.fill_hits(annotation, strandedness, first_read, n_reads)
            return Hits(*[out[k][:n] for k in ("start", "end", "meta", "nh", "read_key")])
        from concurrent.futures import ThreadPoolExecutor
        chunks = []
        step = (n_reads + threads - 1) // threads
        for t in range(threads):
            a = first_read + t * step
            b = min(first_read + n_reads, a + step)
            if a < b:
                chunks.append((a, b - a))
        with ThreadPoolExecutor(threads) as ex:
            counts = list(ex.map(lambda c: self.count_hits(*c), chunks))
            total = sum(counts)
            out = {"start": np.empty(total, np.uint32), "end": np.empty(total, np.uint32), "meta": np.empty(total, np.uint32),
                   "nh": np.empty(total, np.uint32), "read_key": np.empty(total, np.uint64)}
            offs = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int64)
            list(ex.map(lambda co: self.fill_hits(annotation, strandedness, co[0][0], co[0][1], out, int(co[1])), zip(chunks, offs)))
        return Hits(*[out[k] for k in ("start", "end", "meta", "nh", "read_key")])

    def __del__(self):
        if getattr(self, "_h", None):
            lib().mmh_synth_free(self._h)
            self._h = None
