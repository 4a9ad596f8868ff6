#!/usr/bin/env python
"""Turns one evidence pass (tools/gpu_round.sh <tag>) into the tracked files under profiles/:
    python tools/make_profiles.py <tag> [<out-prefix>]
  profiles/<prefix>_launches_{C1,C2}.csv   ncu launch lists (gpu__time_duration.sum per launch)
  profiles/<prefix>_ncu_full_{C1,C2}.md    per-kernel summary of the `ncu --set full` captures: traffic, pipe utilisation,
                                           warp-stall breakdown (cycles a warp waits per instruction it issues, by reason)
  profiles/traffic.json                    DRAM bytes per launch of every stage (bench.py: roofline.traffic; C1 keys plain, C2_ prefixed)
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
for arm in ("ours", "ref"):
    src = os.path.join(G, "bench_%s_%s.json" % (arm, tag))
    if os.path.exists(src):
        shutil.copy(src, os.path.join(P, "%s_bench_%s.json" % (prefix, arm)))
SHAPES = {"C1": "C1 (640x480, P=100k, R=1.08M)", "C2": "C2 (1200x680, P=500k, R=7.4M per view)"}
WL = {"C1": "C1_tum_tracking", "C2": "C2_replica_mapping"}
key = {"preprocess_forward_kernel": "preprocess", "render_forward_kernel": "render_forward",
       "render_backward_kernel": "render_backward", "preprocess_backward_kernel": "preprocess_backward"}
pipes = ["smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
         "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
STALLS = ["long_scoreboard", "short_scoreboard", "barrier", "wait", "mio_throttle", "lg_throttle", "math_pipe_throttle", "not_selected",
          "selected", "dispatch_stall", "branch_resolving", "no_instruction", "sleeping", "membar", "drain", "tex_throttle", "misc"]
traffic = {}
tpath = os.path.join(P, "traffic.json")
if os.path.exists(tpath):
    traffic = json.load(open(tpath))
for S in ("C1", "C2"):
    rep = os.path.join(G, "prof_%s_%s.ncu-rep" % (S, tag))
    if not os.path.exists(rep):
        continue
    shutil.copy(os.path.join(G, "launches_%s_%s.csv" % (S, tag)), os.path.join(P, "%s_launches_%s.csv" % (prefix, S)))
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
    binb = 0
    pre = "" if S == "C1" else S + "_"
    for n, a, b in zip(names, rd, wr):
        if n in key:
            traffic[pre + key[n]] = int(a + b)
        else:
            binb += a + b
    traffic[pre + "binning"] = int(binb)
    icol = col("smsp__issue_active.avg.pct_of_peak_sustained_active")
    for j, n in enumerate(names):      # issue-slot utilisation of the same launches (what actually bounds the compositing kernels)
        if n in key:
            traffic[pre + "issue_pct_" + key[n]] = float(rows[2 + j][icol].replace(",", ""))
    traffic[pre + "_source" if pre else "_source"] = "profiles/%s_ncu_full_%s.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)" % (prefix, S)
    with open(os.path.join(P, "%s_ncu_full_%s.md" % (prefix, S)), "w") as out:
        out.write("# ncu --set full, %s, one un-graphed forward+backward step, B200\n\n" % SHAPES[S])
        out.write("Command (tools/gpu_round.sh): `ncu --set full --clock-control none --import-source on -k 'regex:^(preprocess|render|scatter|"
                  "tile_sort|mark)' -s 10 -c 5 python tools/profile_step.py %s 3`.\n" % WL[S])
        out.write("Times under ncu are serialised, cold-cache single launches; the live CUDA-event stage times are in the bench line.\n\n")
        out.write(summary)
        out.write("\n## pipe utilisation (% of peak while active)\n\n| kernel | issue | ALU | FMA | XU (MUFU/conv) | LSU | threads/inst | "
                  "smem bank conflicts |\n|---|---|---|---|---|---|---|---|\n")
        for j, n in enumerate(names):
            out.write("| %s | %s |\n" % (n, " | ".join(rows[2 + j][col(p)][:8] for p in pipes)))
        out.write("\n## warp stalls: cycles a resident warp spends in each state per instruction it issues "
                  "(`smsp__average_warps_issue_stalled_*_per_issue_active.ratio`; the sum is the warp's mean issue interval, "
                  "`selected` = 1 is the issue itself)\n\n| kernel | total | " + " | ".join(STALLS) + " |\n|---|---|" + "---|" * len(STALLS) + "\n")
        for j, n in enumerate(names):
            v = [float(rows[2 + j][col("smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % st)].replace(",", "") or 0) for st in STALLS]
            out.write("| %s | %.2f | %s |\n" % (n, sum(v), " | ".join("%.2f" % x for x in v)))
        out.write("\n## PC-sampling shares (`smsp__pcsamp_warps_issue_stalled_*`, % of the kernel's samples)\n\n| kernel | samples | "
                  + " | ".join(STALLS) + " |\n|---|---|" + "---|" * len(STALLS) + "\n")
        for j, n in enumerate(names):
            def g(k):
                try:
                    return float(rows[2 + j][col(k)].replace(",", "") or 0)
                except ValueError:
                    return 0.0
            v = [g("smsp__pcsamp_warps_issue_stalled_%s" % (st if st != "no_instruction" else "no_instructions")) for st in STALLS]
            tot = sum(v) or 1.0
            out.write("| %s | %d | %s |\n" % (n, tot, " | ".join("%.1f" % (100 * x / tot) for x in v)))
    print(open(os.path.join(P, "%s_ncu_full_%s.md" % (prefix, S))).read())
json.dump(traffic, open(tpath, "w"), indent=1)
