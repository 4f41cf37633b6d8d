#!/usr/bin/env python3
"""Condense an .ncu-rep (ncu --set full) into a small CSV under profiles/: per captured launch the
duration, DRAM traffic, pipe utilisation and the warp-stall breakdown.

    python tools/ncu_summary.py gpurun_out/foo.ncu-rep profiles/ncu_foo_r01.csv
"""
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "launch__registers_per_thread",
    "launch__block_size",
    "launch__grid_size",
    "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, launches = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    names = [k for k in KEEP if k in col]
    names += sorted(h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"))
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(launches))])
        w.writerow(["Kernel Name", ""] + [r[col["Kernel Name"]] for r in launches])
        for k in names:
            w.writerow([k, units[col[k]]] + [r[col[k]] for r in launches])
    print(f"wrote {out}: {len(launches)} launch(es), {len(names)} metrics")


if __name__ == "__main__":
    main()
