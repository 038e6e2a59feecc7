#!/usr/bin/env python
"""ncu_digest.py <report.ncu-rep> <out.json> [note]: the metrics DESIGN.md quotes (time, DRAM bytes, pipe / issue utilisation,
shared-memory wavefronts and conflicts, occupancy limits, stall reasons per issue) from one `ncu --set full` capture, as a small
JSON digest under profiles/ (the .ncu-rep itself is scratch: tens of MB)."""
import csv
import json
import subprocess
import sys

KEEP = (
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum.pct_of_peak_sustained_elapsed",
    "memory_l1_wavefronts_shared_ideal", "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores",
    "sm__icc_request_hit_rate.pct", "sm__cycles_elapsed.max", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
    "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    names, units, vals = rows[0], rows[1], rows[2]
    d = {}
    for n, u, v in zip(names, units, vals):
        if n in KEEP:
            d[n] = {"value": v, "unit": u}
        elif n == "Kernel Name":
            d["kernel"] = v[:300]
        elif "issue_stalled" in n and n.endswith("per_issue_active.ratio"):
            d.setdefault("stalls_per_issue", {})[n.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")] = float(v)
    if len(sys.argv) > 3:
        d["note"] = sys.argv[3]
    json.dump(d, open(out, "w"), indent=1, sort_keys=True)
    st = d.get("stalls_per_issue", {})
    top = sorted(st.items(), key=lambda kv: -kv[1])[:7]
    g = lambda k: d.get(k, {}).get("value")  # noqa: E731
    print(out, "| us", g("gpu__time_duration.sum"), "| issue%", g("smsp__issue_active.avg.pct_of_peak_sustained_active"), "| fma%",
          g("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"), "| lsu-smem%",
          g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"), "| dram%",
          g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), "| inst", g("smsp__inst_executed.sum"), "| smem wf",
          g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"), "| conflicts", g("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
          "| regs", g("launch__registers_per_thread"), "| warps%", g("sm__warps_active.avg.pct_of_peak_sustained_active"), "| elig",
          g("smsp__warps_eligible.avg.per_cycle_active"), "| icc", g("sm__icc_request_hit_rate.pct"))
    print("   stalls:", ", ".join(f"{k} {v:.2f}" for k, v in top))


if __name__ == "__main__":
    main()
