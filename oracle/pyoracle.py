"""TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT PATH.

ctypes wrapper around oracle/_build/liboracle.so (the plain-C restatement, oracle.c) and a
runner for the compiled reference binaries in oracle/_ref/.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")
REF_DIR = os.path.join(_HERE, "_ref")


class _Params(C.Structure):
    _fields_ = [("strategy", C.c_int32), ("overlap", C.c_float), ("rescue_threshold", C.c_float),
                ("read_stats", C.c_int32), ("n_elements", C.c_uint32), ("elem_line", C.c_void_p),
                ("elem_strand", C.c_void_p), ("elem_vicinity", C.c_void_p), ("rand_seed", C.c_uint32)]


class _Features(C.Structure):
    _fields_ = [("n", C.c_uint32), ("n_chr", C.c_uint32), ("chr", C.c_void_p), ("start", C.c_void_p),
                ("end", C.c_void_p), ("type", C.c_void_p), ("strand", C.c_void_p)]


class _Hits(C.Structure):
    _fields_ = [("n", C.c_uint64), ("start", C.c_void_p), ("end", C.c_void_p), ("meta", C.c_void_p),
                ("nh", C.c_void_p), ("read_key", C.c_void_p)]


class _Result(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("n_hits", "n_reads", "n_unique", "n_ambiguous", "n_multiple", "n_unassigned", "n_rescued", "n_rows")] + \
               [("row_mask", C.POINTER(C.c_uint64)), ("row_value", C.POINTER(C.c_double)), ("hit_mask", C.POINTER(C.c_uint64))]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "oracle/_build/liboracle.so"], cwd=os.path.dirname(_HERE))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        L = C.CDLL(_LIB)
        L.orc_run.argtypes = [C.POINTER(_Params), C.POINTER(_Features), C.POINTER(_Hits), C.c_int, C.POINTER(_Result)]
        L.orc_free.argtypes = [C.POINTER(_Result)]
        L.orc_annotate.argtypes = [C.POINTER(_Params), C.POINTER(_Features), C.POINTER(_Hits), C.c_void_p]
        L.orc_glibc_rand.argtypes = [C.c_uint32, C.c_uint64, C.c_void_p]
        _lib = L
    return _lib


STRATEGIES = {"default": 0, "unique": 1, "random": 2, "ratio": 3}


def _pack(elem_line, elem_strand, elem_vicinity, feats, hits, strategy, overlap, rescue_threshold, read_stats, rand_seed):
    keep = [np.ascontiguousarray(elem_line, np.uint16), np.ascontiguousarray(elem_strand, np.uint8),
            np.ascontiguousarray(elem_vicinity, np.uint8),
            np.ascontiguousarray(feats.chr, np.uint32), np.ascontiguousarray(feats.start, np.uint32),
            np.ascontiguousarray(feats.end, np.uint32), np.ascontiguousarray(feats.type, np.uint8),
            np.ascontiguousarray(feats.strand, np.uint8),
            np.ascontiguousarray(hits.start, np.uint32), np.ascontiguousarray(hits.end, np.uint32),
            np.ascontiguousarray(hits.meta, np.uint32), np.ascontiguousarray(hits.nh, np.uint32),
            np.ascontiguousarray(hits.read_key, np.uint64)]
    p = _Params(STRATEGIES.get(strategy, strategy), float(overlap), float(rescue_threshold), int(read_stats), len(keep[0]),
                keep[0].ctypes.data, keep[1].ctypes.data, keep[2].ctypes.data, rand_seed)
    f = _Features(len(keep[4]), int(feats.n_chr), *[k.ctypes.data for k in keep[3:8]])
    h = _Hits(len(keep[8]), *[k.ctypes.data for k in keep[8:13]])
    return keep, p, f, h


def run(elem_line, elem_strand, elem_vicinity, feats, hits, strategy="default", overlap=-1.0, rescue_threshold=1.0,
        read_stats=False, rand_seed=1, want_hit_masks=False):
    """Returns dict(stats=..., rows={mask: value(double)}, hit_mask=array or None)."""
    keep, p, f, h = _pack(elem_line, elem_strand, elem_vicinity, feats, hits, strategy, overlap, rescue_threshold, read_stats, rand_seed)
    r = _Result()
    if lib().orc_run(C.byref(p), C.byref(f), C.byref(h), int(want_hit_masks), C.byref(r)) != 0:
        raise RuntimeError("oracle failed")
    stats = {k: int(getattr(r, k)) for k in ("n_hits", "n_reads", "n_unique", "n_ambiguous", "n_multiple", "n_unassigned", "n_rescued")}
    rows = {int(r.row_mask[i]): float(r.row_value[i]) for i in range(r.n_rows)}
    hm = None
    if want_hit_masks:
        hm = np.ctypeslib.as_array(r.hit_mask, shape=(max(int(h.n), 1),))[:int(h.n)].copy()
    lib().orc_free(C.byref(r))
    return {"stats": stats, "rows": rows, "hit_mask": hm}


def glibc_rand(seed, n):
    out = np.zeros(n, np.uint32)
    lib().orc_glibc_rand(seed, n, out.ctypes.data)
    return out


# ---------------------------------------------------------------- the compiled reference

def ref_binary(kind="fixed"):
    """Path of oracle/_ref/mmannot_<kind> (asis | fixed | dump), or None if it was not built."""
    p = os.path.join(REF_DIR, "mmannot_" + kind)
    return p if os.path.exists(p) else None


def run_reference(args, kind="fixed", cwd=None, timeout=600):
    """Runs the reference binary; returns (returncode, stdout, stderr) as text."""
    exe = ref_binary(kind)
    if exe is None:
        raise FileNotFoundError("oracle/_ref/mmannot_%s missing: run oracle/build_ref.sh where /root/reference exists" % kind)
    pr = subprocess.run([exe] + list(args), cwd=cwd, capture_output=True, text=True, timeout=timeout)
    return pr.returncode, pr.stdout, pr.stderr


def parse_table(text):
    """-o table -> (column names, {row label: [ints]})."""
    lines = [l for l in text.split("\n") if l]
    if not lines:
        return [], {}
    header = lines[0].split("\t")[1:]
    rows = {}
    for l in lines[1:]:
        parts = l.split("\t")
        rows[parts[0]] = [int(x) for x in parts[1:]]
    return header, rows


def parse_stats(stderr):
    """stderr summary block(s) (mm:1807-1818) -> list of dicts, one per sample."""
    out, cur = [], None
    keys = {"# reads:": "n_reads", "# uniquely mapped reads:": "n_unique", "# multi-mapping rescued reads:": "n_rescued",
            "# hits:": "n_hits", "# ambiguous hits:": "n_ambiguous", "# unassigned hits:": "n_unassigned"}
    for line in stderr.split("\n"):
        s = line.strip()
        if s.startswith("Results for "):
            cur = {}
            out.append(cur)
        elif cur is not None:
            for k, name in keys.items():
                if s.startswith(k):
                    cur[name] = int(s[len(k):].split("(")[0].strip().replace(",", ""))
    return out
