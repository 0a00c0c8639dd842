"""Summarise ncu outputs into the text tables committed under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_X.csv   > profiles/rNN_launch_shares.txt
    python profiles/summarize.py raw gpurun_out/prof_X.ncu-rep        > profiles/rNN_top_kernels.txt
"""
import collections
import csv
import re
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_uniform.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum",
           "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    acc = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ki])[:100]
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
        acc[name][0] += 1
        acc[name][1] += v
    tot = sum(v[1] for v in acc.values())
    print(f"# {path}: {sum(v[0] for v in acc.values())} launches, {tot / 1e3:.2f} ms total (ncu: cold cache, serialised; compare SHARES)")
    print(f"{'share':>7} {'n':>5} {'avg us':>10}  kernel")
    for k, v in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1] / tot * 100:6.2f}% {v[0]:5d} {v[1] / v[0]:10.1f}  {k}")


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")][:110])
        for m in METRICS:
            if m in hdr:
                print(f"   {m:85s} {r[hdr.index(m)]:>16s} {units[hdr.index(m)]}")


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
