"""Turn the scratch captures in gpurun_out/ into the tracked summaries under profiles/.
Usage: python tools/make_profiles.py <tag> <sass_lineinfo.txt>   (tag e.g. r1g -> prof_<tag>_raw.csv, launches_<tag>.csv ...)"""
import collections
import csv
import json
import os
import subprocess
import sys

tag, li = sys.argv[1], sys.argv[2]
G, P = "gpurun_out", "profiles"
KEEP = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__icc_request_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "smsp__average_warp_latency_per_inst_issued.ratio", "sm__cycles_elapsed.max",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum"]
rows = list(csv.reader(open(os.path.join(G, "prof_%s_raw.csv" % tag))))
hdr, units, vals = rows[0], rows[1], rows[2]
with open(os.path.join(P, "%s_ncu_metrics.csv" % tag), "w") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit", "value"])
    for h, u, v in zip(hdr, units, vals):
        if h in KEEP or h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            w.writerow([h, u, v])
d = dict(zip(hdr, vals))
# launch list: per-kernel totals and shares
rows = [r for r in csv.reader(open(os.path.join(G, "launches_%s.csv" % tag))) if len(r) > 5]
h = rows[0]
ik, iv = h.index("Kernel Name"), h.index("Metric Value")
tot = collections.Counter()
cnt = collections.Counter()
for r in rows[1:]:
    if r[h.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    name = r[ik].split("(")[0]
    v = float(r[iv].replace(",", ""))
    unit = r[h.index("Metric Unit")]
    v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "msecond": 1.0, "ms": 1.0, "nsecond": 1e-6, "second": 1e3, "s": 1e3}.get(unit, 1.0)
    tot[name] += v
    cnt[name] += 1
allms = sum(tot.values())
with open(os.path.join(P, "%s_launch_shares.csv" % tag), "w") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "launches", "total_ms", "share_pct"])
    for k, v in tot.most_common():
        w.writerow([k, cnt[k], "%.4f" % v, "%.3f" % (100 * v / allms)])
subprocess.check_call("cp %s/launches_%s.csv %s/%s_launches.csv" % (G, tag, P, tag), shell=True)
out = subprocess.check_output([sys.executable, "tools/ncu_by_line.py", os.path.join(G, "prof_%s_source.csv" % tag), li, "60"]).decode()
open(os.path.join(P, "%s_per_line_and_function.txt" % tag), "w").write(out)
for src, dst in (("bench_%s.json" % tag, "%s_bench.json" % tag), ("bench_%s_ref.json" % tag, "%s_bench_reference_arm.json" % tag)):
    if os.path.exists(os.path.join(G, src)):
        subprocess.check_call(["cp", os.path.join(G, src), os.path.join(P, dst)])
print(json.dumps(dict(duration_ms=d.get("gpu__time_duration.sum"), inst=d.get("smsp__inst_executed.sum"), dram_read=d.get("dram__bytes_read.sum"),
                      dram_write=d.get("dram__bytes_write.sum"), step_kernel_share=100 * tot.get("rg_step_kernel", 0) / allms,
                      issue_active=d.get("smsp__issue_active.avg.pct_of_peak_sustained_active"), l1_hit=d.get("l1tex__t_sector_hit_rate.pct"),
                      icc_hit=d.get("sm__icc_request_hit_rate.pct"), regs=d.get("launch__registers_per_thread"))))
