"""python tools/ncu_summary.py <capture.ncu-rep> <out.json>: the per-kernel numbers the docs quote, from `ncu --page raw --csv`
(first launch of every kernel in the capture)"""
import csv, io, json, subprocess, sys

WANT = {
    "gpu__time_duration.sum": "duration_us",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "lts__t_sectors_op_read.sum": "l2_read_sectors",
    "lts__t_sectors_op_red.sum": "l2_red_sectors",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio": "stall_not_selected",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio": "stall_branch_resolving",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "stall_lg_throttle",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio": "stall_mio_throttle",
}
SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3}

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
out, seen = {"source": sys.argv[1], "command": "ncu --set full --clock-control none --import-source on", "kernels": []}, set()
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    if name in seen:
        continue
    seen.add(name)
    k = {"kernel": name}
    for metric, key in WANT.items():
        if metric in hdr:
            i = hdr.index(metric)
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            k[key] = v * SCALE.get(units[i], 1.0) if key in ("duration_us", "dram_read", "dram_write") else v
    if "dram_read" in k:
        k["dram_bytes_total"] = k["dram_read"] + k.get("dram_write", 0.0)
    out["kernels"].append(k)
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(sys.argv[2], [k["kernel"] for k in out["kernels"]])
