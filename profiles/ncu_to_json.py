"""Summarise an `ncu --set full` report as JSON: one entry per captured launch with the metrics the profiles/*.md tables quote.
Usage: python profiles/ncu_to_json.py REPORT.ncu-rep "capture command line" > profiles/rN_ncu_xxx.json   (reads the report with
`ncu -i REPORT --page raw --csv`; runs anywhere ncu is installed, no GPU needed)"""
import csv
import io
import json
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum",
    "launch__grid_size",
    "launch__block_size",
    "launch__cluster_size",
    "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__cycles_active.avg",
    "sm__cycles_elapsed.max",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum",
    "sass__inst_executed_local_loads",
    "sass__inst_executed_local_stores",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
]


def main():
    rep, capture = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    hdr, units = rows[0], rows[1]
    col = {name: i for i, name in enumerate(hdr)}
    launches = []
    for r in rows[2:]:
        m = {}
        for name in METRICS:
            if name in col and r[col[name]] != "":
                m[name] = {"value": r[col[name]], "unit": units[col[name]]}
        launches.append({"kernel": r[col["Kernel Name"]], "metrics": m})
    print(json.dumps({"report": rep, "capture": capture, "launches": launches}, indent=1))


if __name__ == "__main__":
    main()
