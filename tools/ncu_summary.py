"""Summarise an .ncu-rep into a short text file (key metrics + top stall reasons + hottest SASS lines).
    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep profiles/ncu_x_r01.txt
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "smsp__cycles_elapsed.avg.per_second"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep, dst):
    raw = page(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    lines = ["ncu summary of %s" % rep, "kernel: %s" % vals[hdr.index("Kernel Name")], ""]
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            lines.append("%-78s %s %s" % (k, vals[i], units[i]))
    lines.append("")
    lines.append("issue-stall reasons (warps stalled per issue-active cycle):")
    stalls = [(float(vals[i] or 0), h) for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    for v, h in sorted(stalls, reverse=True)[:8]:
        lines.append("  %-70s %.3f" % (h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""), v))
    src = page(rep, "source")
    if len(src) > 2 and "# Samples" in src[1]:
        h = src[1]
        isamp, isrc = h.index("# Samples"), h.index("Source")
        rows = sorted(src[2:], key=lambda r: -int(r[isamp] or 0))[:12]
        total = sum(int(r[isamp] or 0) for r in src[2:]) or 1
        lines.append("")
        lines.append("hottest SASS instructions (share of %d samples):" % total)
        for r in rows:
            lines.append("  %5.1f%%  %s" % (100.0 * int(r[isamp] or 0) / total, r[isrc].strip()[:100]))
    open(dst, "w").write("\n".join(lines) + "\n")
    print("wrote", dst)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
