"""Aggregate an ncu source-page CSV (SASS view) by CUDA source line.

usage: ncu_by_line.py <src.csv from `ncu -i rep --page source --csv`> <cubin> <kernel substring> [top]
Line info comes from `nvdisasm --print-line-info` of the cubin (compile with -lineinfo)."""
import csv
import re
import subprocess
import sys
from collections import defaultdict

src_csv, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "--print-line-info", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
addr2line = {}
cur = None
inside = False
for ln in dis:
    if ln.startswith("//---------------------"):
        inside = kname in ln and ".text." in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        f = m.group(1).split("/")[-1]
        cur = f"{f}:{m.group(2)}"
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
    if m:
        addr2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
by = defaultdict(lambda: [0, 0, 0, 0, 0, 0])
tot = [0, 0, 0, 0, 0, 0]
base = None
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    a = int(r[col["Address"]], 16) if r[col["Address"]].startswith("0x") else int(r[col["Address"]])
    if base is None:
        base = a
    key = addr2line.get(a - base, ("?", ""))[0]
    def g(name):
        try:
            return int(float(r[col[name]] or 0))
        except Exception:
            return 0
    v = [g("Instructions Executed"), g("# Samples"), g("stall_no_inst"), g("stall_wait"), g("stall_short_sb"), g("L1 Wavefronts Shared")]
    for i in range(6):
        by[key][i] += v[i]
        tot[i] += v[i]
print(f"total: inst {tot[0]:,} samples {tot[1]:,} no_inst {tot[2]:,} wait {tot[3]:,} short_sb {tot[4]:,} smem_wavefronts {tot[5]:,}; static instrs {len(addr2line)}")
print(f"{'line':34s} {'inst%':>6s} {'smp%':>6s} {'noinst':>7s} {'wait':>7s} {'shortsb':>7s} {'smemwf%':>7s}")
for k, v in sorted(by.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{k:34s} {100*v[0]/max(tot[0],1):6.2f} {100*v[1]/max(tot[1],1):6.2f} {v[2]:7d} {v[3]:7d} {v[4]:7d} {100*v[5]/max(tot[5],1):7.2f}")
