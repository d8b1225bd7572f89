#!/usr/bin/env python
"""Turn the scratch ncu outputs of one GPU session (gpurun_out/) into the tracked summaries under profiles/.

    python scripts/ncu_summarise.py TAG [--launches gpurun_out/launches_a.csv] [--rep gpurun_out/prof_a.ncu-rep] [--bench gpurun_out/bench_a.json]

writes profiles/TAG_launches.txt  (per-kernel share of the bench command, from the gpu__time_duration.sum pass)
       profiles/TAG_kbatch.txt    (headline metrics of the dominant kernel from the --set full capture, top source lines, stalls)
       profiles/TAG_traffic.json  (dram bytes of that capture per hit -> bench.py's roofline.traffic)
"""
import argparse, collections, csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sectors_op_read.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
    "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct", "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct",
    "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "launch__shared_mem_per_block_dynamic", "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
]


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def launches_summary(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = None
    agg = collections.OrderedDict()
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = num(d["Metric Value"]) or 0.0
        if d["Metric Unit"] in ("us", "usecond"):
            v *= 1e3
        elif d["Metric Unit"] in ("ms", "msecond"):
            v *= 1e6
        a = agg.setdefault(d["Kernel Name"], [0, 0.0, d["Grid Size"], d["Block Size"]])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values()) or 1.0
    out = ["launch list: ncu --metrics gpu__time_duration.sum --clock-control none (cold, serialised per launch: shares, not absolutes)",
           "source: %s" % os.path.relpath(path, ROOT), "total kernel time %.3f ms over %d launches" % (tot / 1e6, sum(a[0] for a in agg.values())), "",
           "%7s %12s %10s %7s  %-14s %s" % ("count", "total_us", "avg_us", "share", "grid x block", "kernel")]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("%7d %12.1f %10.1f %6.1f%%  %-14s %s" % (a[0], a[1] / 1e3, a[1] / 1e3 / a[0], 100 * a[1] / tot,
                                                           a[2].split(",")[0].strip("( ") + "x" + a[3].split(",")[0].strip("( "), k[:110]))
    return "\n".join(out) + "\n"


def raw_metrics(rep, kernel):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", "regex:" + kernel], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    rows = [r for r in rows if len(r) > 10]
    if len(rows) < 3:
        return None
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--launches", default=os.path.join(ROOT, "gpurun_out", "launches_a.csv"))
    ap.add_argument("--rep", default=os.path.join(ROOT, "gpurun_out", "prof_a.ncu-rep"))
    ap.add_argument("--bench", default=os.path.join(ROOT, "gpurun_out", "bench_a.json"))
    ap.add_argument("--kernel", default="k_batch")
    ap.add_argument("--hits", type=float, default=0.0, help="hits processed by the captured launch (for bytes per hit)")
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    if os.path.exists(a.launches):
        with open(os.path.join(ROOT, "profiles", a.tag + "_launches.txt"), "w") as f:
            f.write(launches_summary(a.launches))
    if os.path.exists(a.rep):
        m = raw_metrics(a.rep, a.kernel)
        out = ["ncu --set full --clock-control none --import-source on, one launch of %s" % a.kernel, "source: %s" % os.path.relpath(a.rep, ROOT)]
        if a.note:
            out.append("note: " + a.note)
        out.append("")
        if m:
            out.append("kernel: " + m.get("Kernel Name", ("?", ""))[0][:160])
            for k in METRICS:
                if k in m:
                    out.append("%-78s %18s %s" % (k, m[k][0], m[k][1]))
            rd, wr = num(m["dram__bytes_read.sum"][0]), num(m["dram__bytes_write.sum"][0])
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd *= scale.get(m["dram__bytes_read.sum"][1], 1)
            wr *= scale.get(m["dram__bytes_write.sum"][1], 1)
            traffic = {"kernel": a.kernel, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": rd + wr, "hits_in_launch": a.hits or None,
                       "dram_bytes_per_hit": (rd + wr) / a.hits if a.hits else None, "source": os.path.relpath(a.rep, ROOT), "tag": a.tag}
            json.dump(traffic, open(os.path.join(ROOT, "profiles", a.tag + "_traffic.json"), "w"), indent=1)
            out.append("")
            out.append("dram traffic of the launch: %.1f MB read + %.1f MB written%s" % (
                rd / 1e6, wr / 1e6, (" = %.2f B/hit over %d hits (algorithmic: see DESIGN.md)" % ((rd + wr) / a.hits, a.hits)) if a.hits else ""))
        for script, title in (("ncu_lines.py", "top source lines by warp instructions"), ("ncu_stalls.py", "top source lines by stall samples")):
            r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", script), a.rep, a.kernel, "32"], capture_output=True, text=True)
            out += ["", "---- " + title, r.stdout.rstrip()]
        with open(os.path.join(ROOT, "profiles", a.tag + "_kbatch.txt"), "w") as f:
            f.write("\n".join(out) + "\n")
    if os.path.exists(a.bench):
        txt = open(a.bench).read().strip()
        if txt:
            with open(os.path.join(ROOT, "profiles", a.tag + "_bench.json"), "w") as f:
                f.write(txt + "\n")


if __name__ == "__main__":
    main()
