#!/usr/bin/env python3
"""Condense `ncu --set full` reports into one tracked CSV (profiles/ncu_*.csv): one row per profiled launch, a fixed
list of the counters DESIGN.md / profiles/README.md quote.  The .ncu-rep files themselves stay in gpurun_out/ (scratch).

    python tools/ncu_export.py profiles/ncu_r2.csv gpurun_out/a.ncu-rep [gpurun_out/b.ncu-rep ...]
"""
import csv
import io
import os
import subprocess
import sys

METRICS = [
    ("duration_us", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("regs", "launch__registers_per_thread"),
    ("dyn_smem_B", "launch__shared_mem_per_block_dynamic"),
    ("sm_mhz", "sm__cycles_elapsed.avg.per_second"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("inst_executed", "smsp__inst_executed.sum"),
    ("tensor_pipe_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("tc_pipe_pct", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active"),
    ("tmem_pipe_pct", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active"),
    ("fma_pipe_pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    ("alu_pipe_pct", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
    ("xu_pipe_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    ("lsu_pipe_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("smem_wavefronts_pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
    ("smem_bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    ("dram_read_B", "dram__bytes_read.sum"),
    ("dram_write_B", "dram__bytes_write.sum"),
    ("dram_pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
    ("l2_read_sectors", "lts__t_sectors_op_read.sum"),
    ("l2_write_sectors", "lts__t_sectors_op_write.sum"),
    ("sm_throughput_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
]
STALLS = ["long_scoreboard", "short_scoreboard", "wait", "barrier", "math_pipe_throttle", "mio_throttle", "lg_throttle",
          "not_selected", "selected", "sleeping", "membar", "dispatch_stall", "no_instruction", "branch_resolving",
          "tex_throttle", "drain", "imc_miss", "gmma"]

UNIT_SCALE = {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6,
              "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units = rd[0], rd[1]
    unit = dict(zip(hdr, units))
    for r in rd[2:]:
        yield dict(zip(hdr, r)), unit


def num(d, unit, key):
    v = d.get(key, "")
    if v == "":
        return ""
    try:
        f = float(v.replace(",", ""))
    except ValueError:
        return v
    u = unit.get(key, "")
    if key.startswith("gpu__time"):
        f *= UNIT_SCALE.get(u, 1.0)
    elif key.startswith("dram__bytes") or key.startswith("launch__shared"):
        f *= UNIT_SCALE.get(u, 1.0)
    elif key.endswith("per_second"):
        f *= {"hz": 1e-6, "Khz": 1e-3, "Mhz": 1.0, "Ghz": 1e3}.get(u, 1.0)
    return f"{f:.6g}"


def main():
    dst, reps = sys.argv[1], sys.argv[2:]
    with open(dst, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["report", "kernel"] + [m[0] for m in METRICS] + ["stall_" + s + "_pct" for s in STALLS])
        for rep in reps:
            for d, unit in rows_of(rep):
                row = [os.path.basename(rep), d.get("Kernel Name", "")[:100]]
                row += [num(d, unit, key) for _, key in METRICS]
                row += [num(d, unit, f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio") or
                        num(d, unit, f"smsp__warp_issue_stalled_{s}_per_warp_active.pct") for s in STALLS]
                w.writerow(row)
    print("wrote", dst)


if __name__ == "__main__":
    main()
