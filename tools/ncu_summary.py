"""Condense an .ncu-rep into the handful of numbers DESIGN.md / profiles/ quote.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--grep PATTERN]
"""
import csv
import re
import subprocess
import sys

KEYS = [
    r"^gpu__time_duration\.sum$",
    r"^sm__cycles_elapsed\.max$",
    r"^launch__(grid_size|block_size|registers_per_thread|shared_mem_per_block_dynamic|waves_per_multiprocessor)$",
    r"^dram__bytes_(read|write)\.sum$",
    r"^dram__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^lts__t_bytes\.sum$",
    r"^lts__t_sector_hit_rate\.pct$",
    r"^lts__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum(\.pct_of_peak_sustained_elapsed)?$",
    r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$",
    r"^sm__inst_executed_pipe_tensor.*\.sum$",
    r"^sm__pipe_tensor.*cycles_active.*pct_of_peak_sustained_(active|elapsed)$",
    r"^sm__mem_tensor_cycles_active\.avg\.pct_of_peak_sustained_elapsed$",
    r"^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^smsp__issue_active\.avg\.pct_of_peak_sustained_active$",
    r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$",
    r"^smsp__inst_executed\.sum$",
    r"^smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio$",
    r"^smsp__warp_issue_stalled_.*_per_warp_active\.pct$",
    r"^sm__ops_path_tensor_op_.*\.sum\.pct_of_peak_sustained_elapsed$",
    r"tensor.*pct",
]


def main():
    rep = sys.argv[1]
    extra = sys.argv[3] if len(sys.argv) > 3 and sys.argv[2] == "--grep" else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    pats = [re.compile(k) for k in KEYS] + ([re.compile(extra)] if extra else [])
    for r in rows[2:]:
        name = dict(zip(hdr, r)).get("Kernel Name", "?")
        print(f"== {name[:110]}")
        for h, u, v in zip(hdr, units, r):
            if any(p.search(h) for p in pats):
                try:
                    if float(v.replace(",", "")) == 0 and "tensor" in h:
                        continue
                except ValueError:
                    pass
                print(f"  {h} = {v} {u}")


if __name__ == "__main__":
    main()
