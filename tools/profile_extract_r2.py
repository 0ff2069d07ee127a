"""Round-2 profile summaries: gpurun_out/r2_* (written by tools/gpu_profiles_r2.sh on the B200 box) -> profiles/r2_*.
   python tools/profile_extract_r2.py [capture-stem] [workload text]"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out")
DST = os.path.join(ROOT, "profiles")
stem = sys.argv[1] if len(sys.argv) > 1 else "r2_eventloop_131k"
workload = sys.argv[2] if len(sys.argv) > 2 else "131,584 terminals, PAR, uniform, 10 msgs/terminal: the bench.py workload (30,490,527 events per launch)"

KEEP = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "l1tex__t_sector_hit_rate.pct", "launch__block_size",
        "launch__grid_size", "launch__registers_per_thread", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "sm__icc_request_hit_rate.pct", "sm__icc_requests.sum.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "gcc__cache_requests_type_instruction.sum", "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed",
        "gcc__average_cache_request_hit_rate.pct"]


def to_bytes(v, unit):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return {h: {"value": v, "unit": u} for h, u, v in zip(rows[0], rows[1], rows[-1])}


rep = os.path.join(SRC, stem + ".ncu-rep")
m = raw_page(rep)
kept = {k: m[k] for k in KEEP if k in m}
json.dump({"workload": workload, "command": "ncu --set full --clock-control none --import-source on -k regex:dfd_run_kernel (tools/gpu_profiles_r2.sh)",
           "metrics": kept}, open(os.path.join(DST, stem + "_metrics.json"), "w"), indent=1)
det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True, check=True).stdout
open(os.path.join(DST, stem + "_details.txt"), "w").write(det)
rd = to_bytes(m["dram__bytes_read.sum"]["value"], m["dram__bytes_read.sum"]["unit"])
wr = to_bytes(m["dram__bytes_write.sum"]["value"], m["dram__bytes_write.sum"]["unit"])
print(stem, "dram bytes per launch:", int(rd + wr), "duration", m["gpu__time_duration.sum"])
if stem == "r2_eventloop_131k":
    json.dump({"kernel": "dfd_run_kernel_pool<16>", "workload": workload, "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
               "dram_bytes_per_launch": int(rd + wr), "source": f"ncu --set full --clock-control none, one launch, profiles/{stem}_metrics.json"},
              open(os.path.join(DST, "r2_traffic.json"), "w"), indent=1)
p = os.path.join(SRC, "r2_launches.csv")
if os.path.exists(p):
    txt = open(p).read()
    open(os.path.join(DST, "r2_launches.csv"), "w").write(txt[txt.index('"ID"'):])
