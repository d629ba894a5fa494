"""Aggregate an ncu source-page CSV by solver PHASE: each SASS instruction is attributed to the outermost
line of `file` (default gi_lane.cuh) on its inline chain, then bucketed by the line ranges given.

usage: ncu_by_phase.py <src.csv> <cubin> <kernel substring> <file> name:lo-hi [name:lo-hi ...]"""
import csv, re, subprocess, sys
from collections import defaultdict

INNER = "--inner" in sys.argv   # attribute to the INNERMOST line of `file` on the inline chain (lambda bodies) instead
argv = [a for a in sys.argv if a != "--inner"]
src_csv, cubin, kname, fname = argv[1:5]
ranges = []
for a in argv[5:]:
    nm, r = a.split(":")
    lo, hi = r.split("-")
    ranges.append((nm, int(lo), int(hi)))
dis = subprocess.run(["nvdisasm", "--print-line-info-inline", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
addr2 = {}
inside = False
chain = []
fresh = False
for ln in dis:
    if ln.startswith("//---------------------"):
        inside = kname in ln and ".text." in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if not fresh:
            chain = []
            fresh = True
        chain.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
    if m:
        fresh = False
        outer = [c for c in chain if c[0] == fname]
        inner = chain[0] if chain else ("?", 0)
        addr2[int(m.group(1), 16)] = ((outer[0][1] if INNER else outer[-1][1]) if outer else -1, inner, m.group(2).strip())
rows = list(csv.reader(open(src_csv)))
hi_ = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi_]
col = {h: i for i, h in enumerate(hdr)}
names = ["Instructions Executed", "# Samples", "stall_no_inst", "stall_wait", "stall_short_sb", "stall_long_sb", "stall_branch_resolving", "Thread Instructions Executed"]
by = defaultdict(lambda: [0] * len(names))
byop = defaultdict(lambda: defaultdict(int))
tot = [0] * len(names)
base = None
for r in rows[hi_ + 1:]:
    if len(r) < len(hdr):
        continue
    a = int(r[col["Address"]], 16) if r[col["Address"]].startswith("0x") else int(r[col["Address"]])
    if base is None:
        base = a
    line, inner, sass = addr2.get(a - base, (-1, ("?", 0), ""))
    key = "other"
    for nm, lo, hi in ranges:
        if lo <= line <= hi:
            key = nm
            break
    v = []
    for n in names:
        try:
            v.append(int(float(r[col[n]] or 0)))
        except Exception:
            v.append(0)
    for i in range(len(names)):
        by[key][i] += v[i]
        tot[i] += v[i]
    op = sass.split()[1] if sass.startswith("@") else (sass.split()[0] if sass else "?")
    byop[key][op.split(".")[0]] += v[0]
print("total", dict(zip(names, tot)))
print(f"{'phase':14s} {'inst%':>6s} {'smp%':>6s} {'noinst':>7s} {'wait':>7s} {'shortsb':>7s} {'longsb':>7s} {'branch':>7s} {'lanes':>5s}  top ops")
for k, v in sorted(by.items(), key=lambda kv: -kv[1][1]):
    ops = sorted(byop[k].items(), key=lambda kv: -kv[1])[:6]
    print(f"{k:14s} {100*v[0]/max(tot[0],1):6.2f} {100*v[1]/max(tot[1],1):6.2f} {v[2]:7d} {v[3]:7d} {v[4]:7d} {v[5]:7d} {v[6]:7d} {v[7]/max(v[0],1):5.1f}  "
          + " ".join(f"{o}:{100*c/max(v[0],1):.0f}%" for o, c in ops))
