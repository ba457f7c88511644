#!/usr/bin/env python
"""Summarise one `ncu --set full` report (.ncu-rep) into a small JSON for profiles/.

    python scripts/ncu_summary.py gpurun_out/knrm_tc_full.ncu-rep profiles/r01_v5_knrm_tc_kernel_ncu_full.json [pairs]

Reads the report with `ncu -i ... --page raw --csv` (works without a GPU) and keeps the metrics DESIGN.md / bench.py cite.
With `pairs` (the pairs the captured launch processed) it also prints the DRAM bytes per pair."""
import csv
import json
import subprocess
import sys

KEEP = (
    "Kernel Name", "Block Size", "Grid Size", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subunit_hmma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    pairs = int(sys.argv[3]) if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {}
    for h, u, v in zip(hdr, units, vals):
        if h in KEEP or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            d[h] = f"{v} {u}".strip()
    if pairs:
        def to_bytes(s):
            v, u = s.split()[:2]
            return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        d["pairs_in_launch"] = pairs
        d["dram_bytes_per_pair"] = (to_bytes(d["dram__bytes_read.sum"]) + to_bytes(d["dram__bytes_write.sum"])) / pairs
    json.dump(d, open(out, "w"), indent=1)
    print(json.dumps({k: d[k] for k in d if "stalled" not in k}, indent=1))


if __name__ == "__main__":
    main()
