"""Summarise an .ncu-rep (one `ncu --set full` capture) as JSON: per kernel name, the averages of the counters the roofline
discussion uses.  python profiles/tools/ncu_summary.py <rep> <out.json> [note]"""
import csv
import json
import subprocess
import sys

rep, out_path = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
KEYS = {
    "gpu_time_us": "gpu__time_duration.sum",
    "sm_throughput_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex_throughput_pct": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l2_throughput_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram_throughput_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram_bytes_read": "dram__bytes_read.sum",
    "dram_bytes_write": "dram__bytes_write.sum",
    "l1tex_t_bytes": "l1tex__t_bytes.sum",
    "l2_t_bytes": "lts__t_bytes.sum",
    "warp_instructions": "smsp__inst_executed.sum",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "registers": "launch__registers_per_thread",
    "grid": "launch__grid_size",
    "block": "launch__block_size",
    "waves_per_sm": "launch__waves_per_multiprocessor",
    "smem_wavefronts": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smem_bank_conflicts": "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lsu_data_pipe_pct": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "pipe_alu_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "pipe_fma_pct": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "pipe_fmaheavy_pct": "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "pipe_lsu_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "stall_long_scoreboard": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "stall_short_scoreboard": "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "stall_barrier": "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "stall_wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "stall_math_pipe": "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "stall_mio_throttle": "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "stall_not_selected": "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "stall_branch": "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "stall_no_instruction": "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
}
agg = {}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d["Kernel Name"].split("(")[0].replace("void ", "")
    a = agg.setdefault(name, {"launches": 0})
    a["launches"] += 1
    for k, m in KEYS.items():
        try:
            a[k] = a.get(k, 0.0) + float(d[m].replace(",", ""))
        except (KeyError, ValueError):
            pass
for name, a in agg.items():
    n = a["launches"]
    for k in list(a):
        if k != "launches":
            a[k] = round(a[k] / n, 3)
json.dump({"capture": rep.split("/")[-1], "note": note, "averages_per_launch": agg}, open(out_path, "w"), indent=1)
print(json.dumps(agg, indent=1)[:3000])
