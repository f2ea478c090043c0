"""Text summary of an .ncu-rep (one block per captured kernel launch): what profiles/*.txt hold.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [kernel-name-regex] > profiles/x.txt
"""
import csv
import io
import re
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"), ("launch__shared_mem_per_block_static", "static smem / block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput (% of ncu peak)"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"), ("l1tex__t_sector_hit_rate.pct", "L1 hit rate"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe"),
    ("smsp__pcsamp_sample_count", "pc samples"),
    ("smsp__pcsamp_warps_issue_stalled_long_scoreboard", "  stalled: long scoreboard (global / local loads)"),
    ("smsp__pcsamp_warps_issue_stalled_barrier", "  stalled: barrier"),
    ("smsp__pcsamp_warps_issue_stalled_wait", "  stalled: wait (fixed latency)"),
    ("smsp__pcsamp_warps_issue_stalled_short_scoreboard", "  stalled: short scoreboard (shared memory, MUFU)"),
    ("smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "  stalled: math pipe throttle"),
    ("smsp__pcsamp_warps_issue_stalled_lg_throttle", "  stalled: lg throttle"),
    ("smsp__pcsamp_warps_issue_stalled_not_selected", "  not selected"),
    ("smsp__pcsamp_warps_issue_stalled_selected", "  selected (issuing)"),
]


def main():
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(head)}
    print(f"# {rep}: ncu --set full --clock-control none (per launch; cold-cache, serialised replays)")
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        if pat and not pat.search(name):
            continue
        print(f"\n== {name}")
        for key, label in WANT:
            if key in col and r[col[key]] != "":
                print(f"   {label:52s} {r[col[key]]} {units[col[key]]}")


if __name__ == "__main__":
    main()
