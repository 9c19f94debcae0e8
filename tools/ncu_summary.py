#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU needed): key raw metrics per kernel + instruction / stall-sample share per
CUDA source line.  Usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-substring] > profiles/xxx.txt"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sectors_op_read.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "sm__cycles_elapsed.max"]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    filt = sys.argv[2] if len(sys.argv) > 2 else ""
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    print("== raw metrics (%s) ==" % rep)
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("-- %s" % name)
        for w in WANT:
            if w in hdr:
                print("   %-70s %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                try:
                    if float(r[i]) >= 0.3:
                        print("   %-70s %s" % (h.replace("smsp__average_warps_issue_stalled_", "stall ").replace("_per_issue_active.ratio", ""), r[i]))
                except ValueError:
                    pass
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    kernel, agg = None, {}
    for r in src:
        if not r:
            continue
        if r[0] == "Function Name":
            kernel = r[1]
            continue
        if r[0] in ("File Path", "Line No") or r[0] == "":
            continue
        if filt and filt not in (kernel or ""):
            continue
        try:
            line, inst, samples = int(r[0]), int(r[7]), int(r[6])
        except (ValueError, IndexError):
            continue
        a = agg.setdefault((kernel, line, r[1].strip()[:100]), [0, 0])
        a[0] += inst
        a[1] += samples
    for k in sorted({k[0] for k in agg}):
        items = [(key, v) for key, v in agg.items() if key[0] == k]
        ti, ts = sum(v[0] for _, v in items) or 1, sum(v[1] for _, v in items) or 1
        print("== source lines, %s: %d warp instructions, %d stall samples ==" % (k, ti, ts))
        for key, v in sorted(items, key=lambda kv: -kv[1][0])[:40]:
            print("   line %4d  %5.1f%% inst  %5.1f%% samples   %s" % (key[1], 100.0 * v[0] / ti, 100.0 * v[1] / ts, key[2]))


if __name__ == "__main__":
    main()
