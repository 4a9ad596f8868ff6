#!/usr/bin/env python
"""Condenses an `ncu --set full` report into the per-kernel figures DESIGN.md / profiles/ quote:
duration, DRAM bytes + GB/s, L2 bytes + GB/s, shared-memory wavefronts, global/shared atomic counts,
issue-slot utilisation, registers, occupancy.   python tools/ncu_summary.py report.ncu-rep > profiles/x.md
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
H, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(H)}


def val(r, k, default=None):
    if k not in idx:
        return default
    try:
        return float(r[idx[k]].replace(",", ""))
    except ValueError:
        return default


def scale_bytes(r, k):
    v = val(r, k)
    if v is None:
        return None
    u = units[idx[k]].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def scale_time_us(r, k):
    v = val(r, k)
    u = units[idx[k]].lower()
    return v * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}.get(u, 1)


print("| kernel | time us | DRAM rd MB | DRAM wr MB | DRAM GB/s | L2 MB | L2 GB/s | L2 %peak | L2 red/atom req | smem wavefronts M | gmem atom/red inst | smem atom inst | warp inst M | issue % | regs | occ % |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for r in rows[2:]:
    if len(r) < len(H):
        continue
    name = r[idx["Kernel Name"]].split("(")[0].split("::")[-1]
    t = scale_time_us(r, "gpu__time_duration.sum")
    rd, wr = scale_bytes(r, "dram__bytes_read.sum") or 0, scale_bytes(r, "dram__bytes_write.sum") or 0
    l2 = (val(r, "lts__t_sectors.sum", 0) or 0) * 32.0
    l2red = (val(r, "lts__t_requests_srcunit_tex_op_red.sum", 0) or 0) + (val(r, "lts__t_requests_srcunit_tex_op_atom_dot_alu.sum", 0) or 0)
    l2pct = val(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed", 0) or 0
    sw = val(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", 0) or 0
    ga = (val(r, "smsp__inst_executed_op_global_atom.sum", 0) or 0) + (val(r, "smsp__inst_executed_op_global_red.sum", 0) or 0)
    sa = val(r, "smsp__inst_executed_op_shared_atom.sum", 0) or 0
    wi = val(r, "smsp__inst_executed.sum", 0) or 0
    iss = val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active", 0)
    regs = val(r, "launch__registers_per_thread", 0)
    occ = val(r, "sm__warps_active.avg.pct_of_peak_sustained_active", 0)
    print("| %s | %.1f | %.2f | %.2f | %.0f | %.1f | %.0f | %.0f | %d | %.2f | %d | %d | %.1f | %.1f | %d | %.0f |" % (
        name, t, rd / 1e6, wr / 1e6, (rd + wr) / (t * 1e-6) / 1e9, l2 / 1e6, l2 / (t * 1e-6) / 1e9, l2pct, l2red, sw / 1e6, ga, sa, wi / 1e6,
        iss, regs, occ))
