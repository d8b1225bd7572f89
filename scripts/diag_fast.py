"""Tuning aid (diag build, MMA_DIAG): why hits leave the segment-table fast path, per workload."""
import sys, os, tempfile, ctypes as C
os.environ["MMANNOT_B200_LIB"] = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mmannot_b200/lib/variants/diag.so")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from mmannot_b200 import device
name = sys.argv[1] if len(sys.argv) > 1 else "tair10_srna"
reads = int(sys.argv[2]) if len(sys.argv) > 2 else 4000000
tmp = tempfile.mkdtemp()
wl = bench.Workload(name, tmp)
pinned, n = wl.fill_pinned(0, reads, 8)
for fs in (0, 7):
    ann = device.Annotator(wl.config, strategy="default", overlap=-1.0, max_batch_hits=1 << 25, fast_bin_shift=fs)
    ann.load_features(wl.annotation)
    out = (C.c_ulonglong * 16)()
    device.lib().mma_diag_get(out, 1)
    ann.submit_batch(0, device.HitBatch(n, *[pinned.arrays[k].ctypes.data for k in ("start", "end", "meta", "nh", "read_key")]))
    res = ann.finish(0)
    device.lib().mma_diag_get(out, 1)
    d = list(out)
    names = ["degenerate", "-", "-", "start beyond looked-up segment", "3+ segments", "-", "answer GENERAL", "answered in-segment", "answered cross-segment", "looked up"]
    print("fast_shift", fs, "index bytes", ann.index_bytes(), "hits", n)
    for k, v in zip(names, d):
        print("   %-30s %10d %6.2f%%" % (k, v, 100.0 * v / max(1, d[9])))
    ann.close()
