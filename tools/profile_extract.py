"""Turn the captures of tools/profile_capture.sh (gpurun_out/profile/) into the tracked summaries under profiles/.
   python tools/profile_extract.py r1"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out", "profile")
DST = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"

KEEP = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "l1tex__t_sector_hit_rate.pct", "launch__block_size",
        "launch__grid_size", "launch__occupancy_limit_registers", "launch__registers_per_thread", "lts__t_sector_hit_rate.pct",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "sm__cycles_elapsed.max",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "sm__icc_request_hit_rate.pct"]


def to_bytes(v, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return float(v.replace(",", "")) * scale


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    return {h: {"value": v, "unit": u} for h, u, v in zip(hdr, units, vals)}


def metrics_csv(path):
    """--metrics ... --csv log: one row per metric"""
    txt = open(path).read()
    txt = txt[txt.index('"ID"'):]
    res = {}
    for row in csv.DictReader(io.StringIO(txt)):
        res[row["Metric Name"]] = {"value": row["Metric Value"], "unit": row["Metric Unit"]}
    return res


def main():
    rep = os.path.join(SRC, "window_kernel_1056.ncu-rep")
    m = raw_page(rep)
    kept = {k: m[k] for k in KEEP if k in m}
    name = f"{tag}_window_kernel_1056_m100"
    json.dump(kept, open(os.path.join(DST, name + "_metrics.json"), "w"), indent=1)
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True, check=True).stdout
    open(os.path.join(DST, name + "_details.txt"), "w").write(det)
    rd = to_bytes(m["dram__bytes_read.sum"]["value"], m["dram__bytes_read.sum"]["unit"])
    wr = to_bytes(m["dram__bytes_write.sum"]["value"], m["dram__bytes_write.sum"]["unit"])
    json.dump({"kernel": "dfd_run_kernel", "workload": "dfdally-1056 uniform minimal, 100 msgs/terminal (bench.py workload)",
               "dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "dram_bytes_per_launch": int(rd + wr),
               "source": f"ncu --set full --clock-control none, one launch, profiles/{name}_metrics.json"},
              open(os.path.join(DST, f"{tag}_traffic.json"), "w"), indent=1)
    # launch list
    txt = open(os.path.join(SRC, "launches.csv")).read()
    open(os.path.join(DST, f"{tag}_launches.csv"), "w").write(txt[txt.index('"ID"'):])
    for cfg in ("8320", "1m"):
        p = os.path.join(SRC, f"dram_{cfg}.csv")
        if os.path.exists(p):
            json.dump(metrics_csv(p), open(os.path.join(DST, f"{tag}_window_kernel_{cfg}_dram.json"), "w"), indent=1)
    for f in ("probe_1056.txt", "probe_8320.txt", "probe_1m.txt", "bench_plain.json"):
        p = os.path.join(SRC, f)
        if os.path.exists(p):
            print(f, ":", open(p).read().strip())
    print("dram bytes/launch (bench workload):", int(rd + wr))


main()
