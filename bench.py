#!/usr/bin/env python
"""Benchmark of the read-annotation hot path (BASELINE.json metric: alignment records/s annotated).

A "step" is one pass of the hot path over one whole sample: Counter::clear, every hit batch through
scan + addCount (mma_submit_hits*), the end-of-file flush and the table read-back (mma_finish_sample),
and for N > 1 the cross-GPU merge of the integer tables.

    value     : hits/s with the packed hit buffers already resident in HBM (CUDA events, max over ranks)
    e2e       : hits/s through mma_submit_hits with HOST (page-locked) hit buffers; the host->device copies of
                every batch and the device->host read of the table are inside the timed region
    roofline  : the dominant kernel's algorithmic bytes / its average launch duration (CUDA events on the
                library's own compute stream) against the measured HBM copy peak (MEASURED_PEAKS.json)
    cpu_baseline : the reference itself (oracle/_ref, compiled from /root/reference in the build container)
                timed on this box's host cores on a bounded sample of the same workload

`--impl reference` times only the reference's CPU path (same workload/metric/unit).
Workload: BASELINE.json configs[1] -- synthetic sRNA-Seq single-end reads (NH <= 20) on a TAIR10-shaped
annotation with configTAIR10.txt, 50 M reads per GPU, name-grouped (mapper order).
"""
import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (shape, config key, seed, reads per GPU, read spec, strandedness, reference args)
    "tair10_srna": dict(shape="tair10", config="configTAIR10", seed=20261018 + 2, reads=50_000_000,
                        spec=dict(max_nh=20), strand="F", strategy="default", overlap=-1.0,
                        describe="synthetic sRNA-Seq single-end, NH<=20, TAIR10-shaped GFF3, configTAIR10.txt, -s F, -y default, -l -1"),
    "flybase6_paired": dict(shape="flybase6", config="configFlybase6", seed=20261018 + 5, reads=25_000_000,
                            spec=dict(max_nh=8, paired=True, rna_seq=True, flip_mate2=True), strand="F", strategy="default", overlap=-1.0,
                            describe="synthetic paired-end RNA-Seq (mate 2 strand-flipped: -s FR as -s F), NH<=8, Flybase6-shaped GFF, configFlybase6.txt"),
    "hs38_multi": dict(shape="hs38", config="configHS38", seed=20261018 + 3, reads=20_000_000,
                       spec=dict(max_nh=100), strand="F", strategy="default", overlap=-1.0,
                       describe="synthetic heavy multi-mapping reads, NH<=100, GRCh38-shaped GTF, configHS38.txt"),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


class Workload:
    """Config + synthetic annotation (files in a scratch directory) of one benchmark shape."""

    def __init__(self, name, tmp, gene_scale=1.0):
        from mmannot_b200 import host
        self.w = WORKLOADS[name]
        self.name = name
        self.tmp = tmp
        cfgs = json.load(open(os.path.join(ROOT, "tests", "golden", "configs.json")))
        self.config_path = os.path.join(tmp, self.w["config"] + ".txt")
        with open(self.config_path, "w") as f:
            f.write(cfgs[self.w["config"]])
        self.synth = host.Synth(self.w["shape"], self.w["seed"], gene_scale=gene_scale, **self.w["spec"])
        self.gtf_path = os.path.join(tmp, "annotation.gff")
        self.synth.write_annotation(self.gtf_path)
        self.config = host.Config(self.config_path)
        self.annotation = host.Annotation(self.config, self.gtf_path)
        self.first_read = 0

    def fill_pinned(self, first_read, n_reads, threads):
        """Packed hits of the read range, generated straight into page-locked buffers."""
        from concurrent.futures import ThreadPoolExecutor
        from mmannot_b200 import device
        threads = max(1, min(threads, n_reads // 1000 or 1))
        step = (n_reads + threads - 1) // threads
        chunks = [(first_read + t * step, min(step, n_reads - t * step)) for t in range(threads) if t * step < n_reads]
        with ThreadPoolExecutor(threads) as ex:
            counts = list(ex.map(lambda c: self.synth.count_hits(*c), chunks))
            total = int(sum(counts))
            pinned = device.PinnedHits(max(total, 1))
            offs = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int64)
            out = pinned.arrays
            got = list(ex.map(lambda co: self.synth.fill_hits(self.annotation, self.w["strand"], co[0][0], co[0][1], out, int(co[1]))[1],
                              zip(chunks, offs)))
        assert [int(g) for g in got] == [int(c) for c in counts]
        return pinned, total


def bind_to_gpu_numa_node(gpu_index):
    """Runs this rank on the CPUs of the NUMA node its GPU hangs off, BEFORE the page-locked buffers are allocated, so that
    they are local to the GPU's PCIe root (with 8 ranks on a two-socket box half of the host->device traffic would
    otherwise cross the socket interconnect).  Best effort: silently does nothing when the topology cannot be read."""
    try:
        exe = shutil.which("nvidia-smi")
        out = subprocess.run([exe, "-i", str(gpu_index), "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True, timeout=20).stdout.strip()
        bus = out.lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]  # sysfs uses a 4-digit PCI domain
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:  # noqa: BLE001
        pass
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        exe = shutil.which("nvidia-smi")
        if not exe:
            return
        fd, self.path = tempfile.mkstemp(suffix=".clocks.csv")
        os.close(fd)
        self.proc = subprocess.Popen([exe, "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                     stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for k, nme in enumerate(names):
                if p[5 + k].lower().startswith("active"):
                    reasons.add(nme)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------ reference arm

def reference_binary():
    from oracle import pyoracle
    return pyoracle.ref_binary("fixed")


def time_reference(wl, n_files, reads_per_file, first_read=0, reps=1, warmup=0):
    """The reference's own CPU path (oracle/_ref/mmannot_fixed: unmodified mmannot.cpp + the one-line setFlags repair
    without which its strand is undefined) on `n_files` BAM files of the workload, -t n_files (one thread per
    BAM is all it can use).  Returns (records/s list per rep, records, cores used, description)."""
    exe = reference_binary()
    if exe is None:
        raise FileNotFoundError("oracle/_ref/mmannot_fixed missing (built by oracle/build_ref.sh where /root/reference exists)")
    from concurrent.futures import ThreadPoolExecutor
    bams = [os.path.join(wl.tmp, "ref_%d.bam" % i) for i in range(n_files)]
    with ThreadPoolExecutor(n_files) as ex:
        list(ex.map(lambda i: wl.synth.write_bam(bams[i], first_read + i * reads_per_file, reads_per_file), range(n_files)))
        records = sum(ex.map(lambda i: wl.synth.count_hits(first_read + i * reads_per_file, reads_per_file), range(n_files)))
    empty = os.path.join(wl.tmp, "ref_empty.bam")
    wl.synth.write_bam(empty, 0, 0)
    w = wl.w
    base = [exe, "-a", wl.gtf_path, "-c", wl.config_path, "-s", w["strand"], "-y", w["strategy"], "-l", repr(w["overlap"]), "-o", os.devnull]

    def run(files, threads):
        t0 = time.perf_counter()
        pr = subprocess.run(base + ["-r"] + files + ["-t", str(threads)], capture_output=True, text=True)
        return time.perf_counter() - t0, pr

    # fixed cost of a run (config + annotation load), subtracted: at the full size of the workload it is negligible
    t_load = min(run([empty], 1)[0] for _ in range(2))
    out = []
    mode = "-t %d in one process" % n_files
    for it in range(warmup + reps):
        dt, pr = run(bams, n_files)
        if pr.returncode != 0:  # the reference's -t > 1 path has unsynchronised table updates (mmannot.cpp:2136)
            mode = "%d independent -t 1 processes" % n_files
            t0 = time.perf_counter()
            procs = [subprocess.Popen(base + ["-r", b, "-t", "1"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for b in bams]
            rcs = [p.wait() for p in procs]
            dt = time.perf_counter() - t0
            if any(rcs):
                raise RuntimeError("reference run failed: " + pr.stderr[-500:])
        if it >= warmup:
            out.append(records / max(dt - t_load, 1e-9))
    for b in bams + [empty]:
        os.unlink(b)
    desc = "%d BAM x %d reads (%d records) of the workload, reads %d.., %s, wall minus %.2fs annotation load" % (
        n_files, reads_per_file, records, first_read, mode, t_load)
    return out, records, n_files, desc


def scratch_root():
    """Where the synthetic GFF / BAM files of BOTH arms go: a RAM-backed tmpfs when the box has one with room (the files were
    written a moment ago, so either way they are read from memory -- but on a disk-backed /tmp the write-back of a freshly written
    0.8 GB BAM runs while it is read, and file -> table time then measures the box's disk: 260 ms on one box, 1600 ms on another)."""
    try:
        st = os.statvfs("/dev/shm")
        if st.f_bavail * st.f_frsize >= (12 << 30) and os.access("/dev/shm", os.W_OK):
            return "/dev/shm"
    except OSError:
        pass
    return None


def run_reference_arm(args):
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    tmp = tempfile.mkdtemp(prefix="mmannot_bench_", dir=scratch_root())
    try:
        wl = Workload(args.workload, tmp)
        cores = os.cpu_count() or 1
        n_files = args.ref_threads or max(1, min(cores, 32))
        vals, records, used, desc = time_reference(wl, n_files, args.ref_reads, reps=args.steps, warmup=args.warmup)
        v = statistics.median(vals)
        line = {"impl": "reference", "metric": "alignment_records_per_sec", "value": v, "unit": "records/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * records / v, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
                "config": bench_config(wl, args, max(1, args.gpus)), "config_detail": {"sample_reads_per_file": args.ref_reads, "files": n_files},
                "cpu_baseline": {"value": v, "unit": "records/s", "cores": used, "kind": "reference", "sample": desc},
                "e2e": {"value": v, "unit": "records/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0, "host_cores": cores}
        emit(line)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return 0


# ------------------------------------------------------------------------------------------ product arm

class HitSet:
    """One rank's hits of a workload: page-locked host arrays, a device-resident copy, and the batch descriptors of both."""
    ISZ = {"start": 4, "end": 4, "meta": 4, "nh": 4, "read_key": 8}
    KEYS = ("start", "end", "meta", "nh", "read_key")

    def __init__(self, pinned, n_hits, dev, own_pinned=True):
        import torch
        self.pinned, self.n, self.own = pinned, int(n_hits), own_pinned
        self.dev = {k: torch.from_numpy(pinned.arrays[k][:max(self.n, 1)].view(np.int32 if k != "read_key" else np.int64)).to(dev)
                    for k in self.KEYS}
        torch.cuda.synchronize()

    def batches(self, where, size):
        from mmannot_b200 import device
        ptr = (lambda k: self.dev[k].data_ptr()) if where == "device" else (lambda k: self.pinned.arrays[k].ctypes.data)
        out = []
        for a in range(0, self.n, size):
            n = min(size, self.n - a)
            out.append(device.HitBatch(n, *[ptr(k) + a * self.ISZ[k] for k in self.KEYS]))
        return out

    def close(self):
        self.dev = {}
        if self.own:
            self.pinned.close()


class SortedView:
    """The first `n` hits of a PinnedHits in another record order (page-locked copy)."""

    def __init__(self, src, n, order):
        from mmannot_b200 import device
        self.inner = device.PinnedHits(max(n, 1))
        for k in HitSet.KEYS:
            self.inner.arrays[k][:n] = src.arrays[k][:n][order]
        self.arrays = self.inner.arrays

    def close(self):
        self.inner.close()


def roofline_of(tm, n_hits, steps, peak, peak_src, kernel):
    """Algorithmic bytes of one launch of the batch kernel: 24 B per hit read once (start, end, meta, nh
    u32 + read key u64); the 16 B/feature index and the 8 B/row table are per sample, not per launch."""
    n_batches = max(1, int(tm["batches"]))
    bytes_per_launch = 24.0 * n_hits * steps / n_batches
    avg_ms = tm["ms_batch"] / n_batches
    achieved = bytes_per_launch / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
            "traffic_source": None, "kernel": kernel, "bytes_per_launch": bytes_per_launch, "avg_launch_ms": avg_ms,
            "launches_timed": n_batches, "peak_source": peak_src,
            "kernel_ms_per_step": {k: tm[k] / steps for k in ("ms_batch", "ms_close", "ms_finish")},
            "segment_table_miss_frac": tm["fast_miss"] / max(1, n_hits)}


def run_product_arm(args):
    import torch
    import torch.distributed as dist

    from mmannot_b200 import device, multi

    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    numa_node = bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cores = os.cpu_count() or 1
    threads = max(1, min(64, cores // max(1, world), len(os.sched_getaffinity(0))))
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, stream):
        """K passes bracketed by barrier + synchronize; CUDA events on the library's own compute stream, max over ranks."""
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        t0 = time.perf_counter()
        ev[0].record(stream)
        res = None
        for _ in range(steps):
            res = fn()
        ev[1].record(stream)
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        ms = max(ev[0].elapsed_time(ev[1]), 0.0)
        if world > 1:
            t = torch.tensor([ms, wall_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall_ms = float(t[0]), float(t[1])
        return ms, wall_ms, res

    def all_sum(v):
        if world == 1:
            return int(v)
        t = torch.tensor([int(v)], device=dev, dtype=torch.int64)
        dist.all_reduce(t)
        return int(t[0])

    same = lambda x, y: np.array_equal(x[0], y[0]) and np.array_equal(device.sort_rows(x[1]), device.sort_rows(y[1]))

    def check_merge(ann, step_local, merged, what):
        """N > 1, outside the timed region: the table the GPUs merged (export / all-gather / import) must be the sum of the
        per-rank tables, formed here on the host from every rank's own mma_finish_sample."""
        if world == 1:
            return True
        loc = step_local()
        gathered = [None] * world
        dist.all_gather_object(gathered, (loc[0], loc[1]))
        stats = np.sum([g[0] for g in gathered], axis=0)
        acc = {}
        for g in gathered:
            u = g[1].view(np.uint64)
            for i in range(len(u)):
                key = (int(u[i, 0]), int(u[i, 1]))
                acc[key] = acc.get(key, 0) + int(u[i, 2])
        got = merged[1].view(np.uint64)
        got_d = {(int(got[i, 0]), int(got[i, 1])): int(got[i, 2]) for i in range(len(got))}
        acc = {k: v for k, v in acc.items() if v}
        assert np.array_equal(stats, merged[0]), "%s: merged counters %s != sum of the per-rank counters %s" % (what, merged[0], stats)
        assert got_d == acc, "%s: device-merged table differs from the host-side sum of the per-rank tables" % what
        return True

    def check_oracle(wl, ann, hs, n_reads, per_name):
        """This rank's table on its first reads against the CPU oracle (test infrastructure, here as the checker only)."""
        from oracle import pyoracle
        from mmannot_b200 import host
        first = wl.first_read
        n = wl.synth.count_hits(first, n_reads)
        h = host.Hits(*[np.array(hs.pinned.arrays[k][:n]) for k in HitSet.KEYS])
        ref = pyoracle.run(wl.config.elem_line, wl.config.elem_strand, wl.config.elem_vicinity, wl.annotation, h,
                           strategy=wl.w["strategy"], overlap=wl.w["overlap"])
        ann.reset(0)
        ann.submit(0, h)
        res = ann.finish(0)
        got = device.values_by_mask(res["rows"])
        assert set(got) == set(ref["rows"]), "oracle check: element sets differ on the first %d reads" % n_reads
        for m, v in ref["rows"].items():
            assert abs(got[m] - v) <= 1e-9 * max(1.0, abs(v)), "oracle check: count of set %x" % m
        assert res["stats"] == ref["stats"], "oracle check: counters %s != %s" % (res["stats"], ref["stats"])
        return n

    def device_resident(wl, hs, strategy=None, merge=True, steps=None, dev_batch=None, what=""):
        """Device-resident passes of one workload: -> dict for the `workloads` key (and the objects the main workload needs)."""
        w = wl.w
        steps = steps or args.steps
        dev_batch = dev_batch or max(args.device_batch, args.batch)
        strategy = strategy or w["strategy"]
        ann = device.Annotator(wl.config, strategy=strategy, overlap=w["overlap"], n_samples=1, max_batch_hits=dev_batch,
                               device=local_rank, table_log2=args.table_log2, fast_bin_shift=args.fast_shift, bin_shift=args.bin_shift)
        ann.load_features(wl.annotation)
        stream = torch.cuda.ExternalStream(ann.stream_ptr(), device=dev)
        dev_batches = hs.batches("device", dev_batch)
        merged = world > 1 and merge

        def step_local():
            ann.reset(0)
            for b in dev_batches:
                ann.submit_device(0, b)
            return ann.finish_arrays(0, sort=False)

        def step():
            if not merged:
                return step_local()
            ann.reset(0)
            for b in dev_batches:
                ann.submit_device(0, b)
            return multi.merge_on_device(ann, 0, dev)

        for _ in range(max(2, min(args.warmup, 3))):
            step()
        ann.timing_enable(True)
        ann.timing_reset()
        ms, wall, res = timed(step, steps, stream)
        tm = ann.timing()
        ann.timing_enable(False)
        total_hits = all_sum(hs.n)
        stats = {k: int(v) for k, v in zip(multi.STAT_KEYS, res[0])}
        if not merged and world > 1:  # one sample per GPU: the counters of the samples, summed for the invariants below
            t = torch.tensor(res[0], device=dev, dtype=torch.int64)
            dist.all_reduce(t)
            stats = {k: int(v) for k, v in zip(multi.STAT_KEYS, t.cpu().numpy())}
        # size-independent properties of the synthetic workloads: every read is complete (NH records each), so the reads counted
        # must be the reads generated and the hits counted the hits submitted, over all ranks
        checks = []
        if strategy in ("default", "ratio"):
            assert stats["n_hits"] == total_hits, "%s: hits counted %d != hits submitted %d" % (what, stats["n_hits"], total_hits)
            checks.append("hits counted == hits submitted")
        if strategy == "default" and hs.complete_reads:
            per_name = 2 if w["spec"].get("paired") else 1  # both mates carry the name and the NH: two countdowns per name
            want = per_name * hs.complete_reads * world
            assert stats["n_reads"] == want, "%s: reads counted %d != reads generated %d" % (what, stats["n_reads"], want)
            checks.append("reads counted == reads generated")
        if merged:
            check_merge(ann, step_local, res, what)
            checks.append("device-merged table == host-side sum of the per-rank tables")
        ms_per_step = ms / steps
        out = {"value": total_hits / (ms_per_step * 1e-3), "unit": "records/s", "ms_per_step": ms_per_step, "hits_per_gpu": hs.n,
               "reads_per_gpu": hs.reads, "features": int(wl.annotation.n), "elements": int(wl.config.n_elements), "strategy": strategy,
               "describe": w["describe"], "roofline": roofline_of(tm, hs.n, steps, peak, peak_src, ann.dominant_kernel()),
               "gpu_launches": int(tm["launches"]), "table_rows": int(len(res[1])), "stats": stats, "checks": checks,
               "multi_gpu": ("read-name ranges, index replicated, tables merged on the GPUs around one NCCL all-gather" if merged
                             else "one sample per GPU, no exchange") if world > 1 else "single GPU"}
        return out, ann, res, stream, wall / steps, tm

    def make_hitset(wl, reads):
        """Read-name-range sharding: rank r owns reads [r * reads, (r + 1) * reads) -- every record of a read stays on one GPU."""
        t0 = time.time()
        wl.first_read = rank * reads
        pinned, n_hits = wl.fill_pinned(rank * reads, reads, threads)
        hs = HitSet(pinned, n_hits, dev)
        hs.reads, hs.complete_reads = reads, reads
        log("[rank %d] workload %s: %d features, %d reads -> %d hits generated in %.1fs (%d threads)" % (
            rank, wl.name, wl.annotation.n, reads, n_hits, time.time() - t0, threads))
        return hs

    tmp = tempfile.mkdtemp(prefix="mmannot_bench_%d_" % rank, dir=scratch_root())
    try:
        wl = Workload(args.workload, tmp)
        w = wl.w
        reads = args.reads or w["reads"]
        hs = make_hitset(wl, reads)
        n_hits = hs.n
        batch = args.batch                                  # hits per call, host-buffer passes (copies overlap the kernels)
        dev_batch = max(args.device_batch, batch)           # hits per call, device-resident pass
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        main, ann, res_dev, stream, wall_dev, tm = device_resident(wl, hs, what=args.workload)
        index_bytes, segments = ann.index_bytes(), ann.index_segments()
        # ---- the same sample through the C ABI from HOST buffers (page-locked), copies and table read-back inside the timed
        #      region: in the compact transfer format the host decoder emits, and in the wide arrays of mma_submit_hits
        host_batches = hs.batches("host", batch)
        packed_batches = []
        if not args.e2e_wide:
            for a in range(0, n_hits, batch):
                n = min(batch, n_hits - a)
                packed_batches.append(device.PackedHits(*[hs.pinned.arrays[k][a:a + n] for k in HitSet.KEYS]))

        def finish():
            return multi.merge_on_device(ann, 0, dev) if world > 1 else ann.finish_arrays(0, sort=False)

        def step_e2e():
            ann.reset(0)
            if packed_batches:
                for pb in packed_batches:
                    ann.submit_packed(0, pb.batch)
            else:
                for b in host_batches:
                    ann.submit_batch(0, b)
            return finish()

        def step_e2e_wide():
            ann.reset(0)
            for b in host_batches:
                ann.submit_batch(0, b)
            return finish()

        for _ in range(min(args.warmup, 2)):
            step_e2e()
        _, wall_e2e, res_e2e = timed(step_e2e, args.steps, stream)
        _, wall_e2e_wide, res_e2e_wide = timed(step_e2e_wide, max(1, args.steps // 2), stream)
        clocks = sampler.stop() if rank == 0 else None
        assert same(res_e2e_wide, res_e2e), "packed and wide host-buffer passes disagree"
        assert same(res_dev, res_e2e), "device-resident and host-buffer passes disagree"
        main["checks"].append("device-resident pass == compact host pass == wide host pass")
        if args.oracle_reads:
            n_checked = check_oracle(wl, ann, hs, min(args.oracle_reads, reads), 1)
            main["checks"].append("table of this rank's first %d reads (%d hits) == CPU oracle" % (min(args.oracle_reads, reads), n_checked))
        total_hits = all_sum(n_hits)
        e2e_value = total_hits / (wall_e2e / args.steps * 1e-3)
        h2d = sum(pb.h2d_bytes for pb in packed_batches) if packed_batches else 24 * n_hits
        d2h = ann.table_readback_bytes()
        for pb in packed_batches:
            pb.close()
        packed_batches = []

        # ---- the other shapes BASELINE.json names, device-resident (value, ms per step, roofline of the batch kernel)
        workloads = {}
        if not args.no_secondary:
            # config 4: one sample per GPU, -y ratio (-e 80 is inert without -m): this rank's sample is its read range
            r4, a4, _, _, _, _ = device_resident(wl, hs, strategy="ratio", merge=False, what="ratio_samples")
            a4.close()
            r4["describe"] = "BASELINE config 4: one synthetic sRNA sample per GPU (TAIR10 shape), -y ratio -e 80"
            workloads["ratio_%dsamples" % world] = r4
            # coordinate-sorted variant of config 2 (deferred path: every multi-mapping read is resolved after a sort by name)
            n_cs = wl.synth.count_hits(wl.first_read, min(reads, args.coordsorted_reads))
            order = np.lexsort((hs.pinned.arrays["start"][:n_cs], hs.pinned.arrays["meta"][:n_cs] & 0xFFFFFF))
            sv = SortedView(hs.pinned, n_cs, order)
            hcs = HitSet(sv, n_cs, dev)
            hcs.reads, hcs.complete_reads = min(reads, args.coordsorted_reads), min(reads, args.coordsorted_reads)
            rcs, acs, _, _, _, _ = device_resident(wl, hcs, what="tair10_coordsorted")
            acs.close()
            hcs.close()
            rcs["describe"] = "config 2 hits in coordinate-sorted order (deferred path)"
            workloads["tair10_coordsorted"] = rcs
        ann.close()
        hs.close()
        if not args.no_secondary:
            for name in ("flybase6_paired", "hs38_multi"):
                if name == args.workload:
                    continue
                tmp2 = tempfile.mkdtemp(prefix="mmannot_bench_%d_%s_" % (rank, name), dir=scratch_root())
                try:
                    wl2 = Workload(name, tmp2)
                    hs2 = make_hitset(wl2, args.secondary_reads or wl2.w["reads"])
                    r2, a2, _, _, _, _ = device_resident(wl2, hs2, what=name)
                    if args.oracle_reads:
                        nck = check_oracle(wl2, a2, hs2, min(args.oracle_reads, hs2.reads), 1)
                        r2["checks"].append("table of this rank's first %d reads (%d hits) == CPU oracle" % (min(args.oracle_reads, hs2.reads), nck))
                    a2.close()
                    hs2.close()
                    workloads[name] = r2
                finally:
                    shutil.rmtree(tmp2, ignore_errors=True)

        if rank == 0:
            roofline = main["roofline"]
            # dram bytes of the dominant kernel per launch, from the committed ncu --set full capture (bytes per hit there x
            # the hits of an average launch here); null when no capture of this kernel is on file
            try:
                import glob
                for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")), reverse=True):
                    tj = json.load(open(path))
                    if tj.get("kernel") == roofline["kernel"] and tj.get("dram_bytes_per_hit"):
                        roofline["traffic"] = tj["dram_bytes_per_hit"] * n_hits * args.steps / roofline["launches_timed"]
                        roofline["traffic_source"] = os.path.relpath(path, ROOT)
                        break
            except Exception:  # noqa: BLE001
                pass
            cpu = None
            if not args.no_cpu_baseline and world == 1:  # (the contract: rank 0 at N = 1 only)
                try:
                    n_files = args.ref_threads or 1
                    vals, records, used, desc = time_reference(wl, n_files, args.cpu_reads, reps=1)
                    cpu = {"value": vals[0], "unit": "records/s", "cores": used, "kind": "reference", "sample": desc}
                except Exception as e:  # noqa: BLE001
                    cpu = {"value": None, "unit": "records/s", "cores": 0, "kind": "reference", "sample": "unavailable: %s" % e}
            e2e_file = None
            if not args.no_file and world == 1:  # (one file through one GPU: nothing to add at N > 1)
                try:
                    e2e_file = time_cli_file(wl, args.file_reads, threads)
                except Exception as e:  # noqa: BLE001
                    e2e_file = {"value": None, "error": str(e)}
            line = {"metric": "alignment_records_per_sec", "value": main["value"], "unit": "records/s", "n_gpus": world, "steps": args.steps,
                    "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                    "dtype": "u32", "data": "synthetic",
                    "config": bench_config(wl, args, world),
                    "config_detail": {"hits_per_gpu": n_hits, "features": int(wl.annotation.n), "elements": int(wl.config.n_elements), "batch_hits": dev_batch,
                                      "batch_hits_e2e": batch, "index_bytes": index_bytes, "segments": segments, "gb_per_pass": 24e-9 * n_hits},
                    "e2e": {"value": e2e_value, "unit": "records/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                            "ms_per_step": wall_e2e / args.steps, "timer": "host wall clock around synchronize",
                            "format": "compact transfer format as the host decoder emits it (mma_submit_hits_packed: 8 B/hit + 8 B/run, expanded on the device); "
                                      "the buffers hold decoded hits -- BAM inflate and record parsing are NOT in this number, see e2e_file"
                                      if not args.e2e_wide else "wide (24 B/hit)",
                            "wide_format_value": total_hits / (wall_e2e_wide / max(1, args.steps // 2) * 1e-3), "wide_h2d_bytes_per_step": 24 * n_hits},
                    "e2e_file": e2e_file,
                    "gpu_launches": main["gpu_launches"],
                    "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks, "host_cores": cores, "numa_node": numa_node,
                    "wall_ms_per_step_device_resident": wall_dev,
                    "stats": main["stats"], "table_rows": main["table_rows"], "checks": main["checks"], "workloads": workloads}
            emit(line)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return 0


def bench_config(wl, args, world):
    """The `config` object, IDENTICAL in both arms (the driver compares them); what only one arm knows goes to `config_detail`."""
    reads = args.reads or wl.w["reads"]
    return {"workload": wl.w["describe"], "name": wl.name, "shape": wl.w["shape"], "strategy": wl.w["strategy"], "strand": wl.w["strand"],
            "overlap": wl.w["overlap"], "reads_per_gpu": reads, "order": "name-grouped (mapper order)",
            "l2": "inputs (24 B per hit, ~2.1 hits per read: GBs per pass) larger than L2; no explicit flush",
            "sharding": "read-name ranges, index replicated, tables merged on the GPUs around one NCCL all-gather" if world > 1 else "single GPU"}


def time_cli_file(wl, n_reads, threads):
    """File -> table through the drop-in command line (BGZF inflate, BAM parse, pack, copies, kernels, table): the number that
    compares like with like with the reference arm.  The BAM is written here (untimed); the command line reports the time of
    its Counter::read."""
    exe = os.path.join(ROOT, "mmannot_b200", "bin", "mmannot_b200")
    bam = os.path.join(wl.tmp, "e2e_file.bam")
    t0 = time.time()
    wl.synth.write_bam_parallel(bam, 0, n_reads, threads)
    records = wl.synth.count_hits(0, n_reads)
    t_write = time.time() - t0
    w = wl.w
    env = dict(os.environ, MMANNOT_B200_TIMING="1")
    cmd = [exe, "-a", wl.gtf_path, "-c", wl.config_path, "-s", w["strand"], "-y", w["strategy"], "-l", repr(w["overlap"]), "-o", os.devnull, "-r", bam]
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        pr = subprocess.run(cmd, capture_output=True, text=True, env=env)
        wall = time.perf_counter() - t0
        if pr.returncode != 0:
            raise RuntimeError("command line failed: " + pr.stderr[-300:])
        read_ms, detail = None, None
        for ln in pr.stderr.splitlines():
            if ln.startswith("[timing] read_ms="):
                read_ms = float(ln.split("=")[1].split()[0])
            elif ln.startswith("[timing] bam file read"):
                detail = ln[len("[timing] "):]
        if read_ms is not None and (best is None or read_ms < best[0]):
            best = (read_ms, wall, detail)
    os.unlink(bam)
    if best is None:
        raise RuntimeError("no timing line from the command line")
    size = None
    return {"value": records / (best[0] * 1e-3), "unit": "records/s", "records": records, "reads": n_reads, "read_ms": best[0],
            "process_wall_s": best[1], "bam_write_s": t_write, "bam_dir": os.path.dirname(bam) + (" (tmpfs)" if bam.startswith("/dev/shm") else ""),
            "breakdown": best[2],
            "what": "one BAM -> count table through mmannot_b200/bin/mmannot_b200, timed inside the process around Counter::read: the compressed "
                    "file over PCIe, BGZF inflate + BAM record parse + batch kernels on the device (mma_submit_bam), table read-back; "
                    "`breakdown` is null when the file took the host decoder instead; process wall time includes annotation load and CUDA start-up"}


_REAL_STDOUT = None


def emit(line):
    """The one JSON line of the contract, on the process's real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # libraries (NCCL prints its version banner) must not write to stdout: everything but the JSON line goes to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="tair10_srna", choices=sorted(WORKLOADS))
    ap.add_argument("--reads", type=int, default=0, help="reads per GPU (default: the workload's full size)")
    ap.add_argument("--batch", type=int, default=1 << 25, help="hits per mma_submit_hits* call of the host-buffer (e2e) passes")
    ap.add_argument("--device-batch", type=int, default=1 << 27, help="hits per mma_submit_hits_device call of the device-resident pass")
    ap.add_argument("--table-log2", type=int, default=0)
    ap.add_argument("--fast-shift", type=int, default=0, help="log2 bin width of the segment answer table (0 = auto, -1 = no table)")
    ap.add_argument("--bin-shift", type=int, default=0)
    ap.add_argument("--cpu-reads", type=int, default=3_000_000, help="reads of the bounded cpu_baseline sample (per BAM)")
    ap.add_argument("--ref-reads", type=int, default=500_000, help="--impl reference: reads per BAM per step")
    ap.add_argument("--ref-threads", type=int, default=0, help="BAM files / threads of the reference run (default: 1 for cpu_baseline, host cores up to 32 for --impl reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the `workloads` entries (other BASELINE shapes)")
    ap.add_argument("--no-file", action="store_true", help="skip e2e_file (BAM -> table through the command line)")
    ap.add_argument("--secondary-reads", type=int, default=0, help="reads per GPU of the secondary workloads (default: their full size)")
    ap.add_argument("--coordsorted-reads", type=int, default=10_000_000, help="reads of the coordinate-sorted variant")
    ap.add_argument("--file-reads", type=int, default=25_000_000, help="reads of the BAM of e2e_file")
    ap.add_argument("--oracle-reads", type=int, default=200_000, help="reads per rank checked against the CPU oracle outside the timed region (0 = off)")
    ap.add_argument("--e2e-wide", action="store_true", help="e2e through the wide 24 B/hit arrays instead of the compact transfer format")
    args = ap.parse_args()
    if args.fast_shift < 0:
        args.fast_shift = None
    if args.warmup < 3:
        log("bench.py: note: fewer than 3 warm-up steps requested")
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_product_arm(args)


if __name__ == "__main__":
    sys.exit(main())
