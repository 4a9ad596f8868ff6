#!/usr/bin/env python
"""Turns one evidence pass (tools/gpu_round.sh <tag>) into the tracked files under profiles/:
    python tools/make_profiles.py <tag> [<out-prefix>]
  profiles/<prefix>_launches_C1.csv   ncu launch list (gpu__time_duration.sum per launch)
  profiles/<prefix>_ncu_full_C1.md    per-kernel summary of the `ncu --set full` capture + pipe utilisation
  profiles/traffic.json               DRAM bytes per launch of every stage (bench.py: roofline.traffic)
  profiles/<prefix>_bench_*.json      the bench lines of the same pass
"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
prefix = sys.argv[2] if len(sys.argv) > 2 else tag
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
shutil.copy(os.path.join(G, "launches_%s.csv" % tag), os.path.join(P, "%s_launches_C1.csv" % prefix))
for arm in ("ours", "ref"):
    src = os.path.join(G, "bench_%s_%s.json" % (arm, tag))
    if os.path.exists(src):
        shutil.copy(src, os.path.join(P, "%s_bench_%s.json" % (prefix, arm)))
rep = os.path.join(G, "prof_%s.ncu-rep" % tag)
summary = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H, U = rows[0], rows[1]
col = H.index
names = [r[col("Kernel Name")].split("(")[0].split("::")[-1].split("<")[0] for r in rows[2:]]


def scaled(k):
    i = col(k)
    m = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(U[i].lower(), 1)
    return [float(r[i].replace(",", "")) * m for r in rows[2:]]


rd, wr = scaled("dram__bytes_read.sum"), scaled("dram__bytes_write.sum")
key = {"preprocess_forward_kernel": "preprocess", "render_forward_kernel": "render_forward",
       "render_backward_kernel": "render_backward", "preprocess_backward_kernel": "preprocess_backward"}
traffic, binb = {}, 0
for n, a, b in zip(names, rd, wr):
    if n in key:
        traffic[key[n]] = int(a + b)
    else:
        binb += a + b
traffic["binning"] = int(binb)
traffic["_source"] = "profiles/%s_ncu_full_C1.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, C1)" % prefix
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
pipes = ["smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
         "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
with open(os.path.join(P, "%s_ncu_full_C1.md" % prefix), "w") as out:
    out.write("# ncu --set full, C1 (640x480, P=100k, R=1.08M), one un-graphed forward+backward step, B200\n\n")
    out.write("Command (tools/gpu_round.sh): `ncu --set full --clock-control none --import-source on -k 'regex:^(preprocess|render|scatter|"
              "tile_sort|mark)' -s 9 -c 4 python tools/profile_step.py C1_tum_tracking 3`.\n")
    out.write("Times under ncu are serialised, cold-cache single launches; the live CUDA-event stage times are in the bench line.\n\n")
    out.write(summary)
    out.write("\n## pipe utilisation (% of peak while active)\n\n| kernel | issue | ALU | FMA | XU (MUFU/conv) | LSU | threads/inst | "
              "smem bank conflicts |\n|---|---|---|---|---|---|---|---|\n")
    for j, n in enumerate(names):
        out.write("| %s | %s |\n" % (n, " | ".join(rows[2 + j][col(p)][:8] for p in pipes)))
print(open(os.path.join(P, "%s_ncu_full_C1.md" % prefix)).read())
