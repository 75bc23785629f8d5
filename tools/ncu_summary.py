"""Condenses ncu captures into the text summaries kept under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_r1.csv          # per-kernel time shares of a --metrics gpu__time_duration.sum list
    python tools/ncu_summary.py raw gpurun_out/mask_r1.ncu-rep [name-regex]  # selected metrics of every launch of a --set full capture
"""
import csv, io, re, subprocess, sys
from collections import defaultdict

WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
    hdr = rows[0]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        name = re.sub(r"^void ", "", r[ki])
        name = re.sub(r"\(.*", "", name)[:100]
        tot[name] += v; cnt[name] += 1
    total = sum(tot.values())
    print(f"# total kernel time {total:.1f} ms over {sum(cnt.values())} launches\n")
    print(f"{'ms':>10}  {'share':>6}  {'launches':>8}  kernel")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{v:10.3f}  {100 * v / total:5.1f}%  {cnt[k]:8d}  {k}")


def raw(path, pat=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        if pat and not re.search(pat, r[ki]):
            continue
        print("== " + r[ki][:110])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"   {w:95s} {r[i]:>18s} {units[i]}")
        print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        raw(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
