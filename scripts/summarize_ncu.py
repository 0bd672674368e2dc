"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.
  python scripts/summarize_ncu.py gpurun_out/launches_r1.csv gpurun_out/prof_r1.ncu-rep r1"""
import collections
import csv
import io
import subprocess
import sys

launch_csv, rep, tag = sys.argv[1], sys.argv[2], sys.argv[3]
lines = [l for l in open(launch_csv) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    k = row["Kernel Name"].split("(")[0].replace("void ", "").replace("dftk::", "").replace(", ", "x")
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v *= {"nsecond": 1e-6, "ns": 1e-6, "usecond": 1e-3, "us": 1e-3, "msecond": 1.0, "ms": 1.0, "second": 1e3}[u]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in agg.values())
out = [f"# ncu launch list ({tag}): gpu__time_duration.sum per kernel, --clock-control none (cold-cache, serialised: compare shares)",
       f"# source: {launch_csv}; command: M=102 REPS=2 ncu --metrics gpu__time_duration.sum -k regex:'k_|kr_' python scripts/profile_probe.py",
       "kernel,launches,total_ms,avg_ms,share"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{k},{v[0]},{v[1]:.3f},{v[1] / v[0]:.3f},{v[1] / tot:.3f}")
open(f"profiles/launches_{tag}.csv", "w").write("\n".join(out) + "\n")

raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]
want = [w for w in want if w in idx]
with open(f"profiles/ncu_full_{tag}.csv", "w") as f:
    f.write(f"# ncu --set full --clock-control none --import-source on ({tag}); one row per captured launch; source {rep}\n")
    f.write("kernel," + ",".join(want) + "\n")
    f.write("unit," + ",".join(units[idx[w]] for w in want) + "\n")
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace(", ", "x")
        f.write(name + "," + ",".join(r[idx[w]].replace(",", "") for w in want) + "\n")
# measured DRAM traffic of the Hpsi-local kernel group per band (bench.py reports it as roofline.traffic): the capture runs
# M_FULL bands in one chunk, one launch of each of the five pipeline kernels
import json
import os
m_full = int(os.environ.get("M_FULL", 51))
group = ("kr_sphere_to_x", "kr_y_backward", "kr_z_apply", "kr_y_forward", "kr_x_to_sphere")
seen, total = set(), 0.0
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("dftk::", "").split("<")[0]
    if name in group and name not in seen:
        seen.add(name)
        def val(col):
            v = float(r[idx[col]].replace(",", ""))
            return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[idx[col]]]
        total += val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
if len(seen) == len(group):
    json.dump(dict(hpsi_local_dram_bytes_per_band=total / m_full, bands_per_launch=m_full, kernels=list(group),
                   source=f"profiles/ncu_full_{tag}.csv (dram__bytes_read.sum + dram__bytes_write.sum of one launch of each of the five "
                          f"kernels at {m_full} bands, ncu --set full --clock-control none)"),
              open("profiles/ncu_traffic.json", "w"), indent=1)
print(open(f"profiles/launches_{tag}.csv").read())
