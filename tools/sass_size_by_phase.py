"""Static SASS size of a kernel by source phase (outermost gi_lane.cuh line on each instruction's inline chain).
usage: sass_size_by_phase.py <cubin> <kernel substring> <file> name:lo-hi ..."""
import re, subprocess, sys
from collections import Counter
cubin, kname, fname = sys.argv[1:4]
ranges = []
for a in sys.argv[4:]:
    nm, r = a.split(":"); lo, hi = r.split("-"); ranges.append((nm, int(lo), int(hi)))
dis = subprocess.run(["nvdisasm", "--print-line-info-inline", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
inside = False; chain = []; fresh = False; cnt = Counter(); tot = 0
for ln in dis:
    if ln.startswith("//---------------------"):
        inside = kname in ln and ".text." in ln; continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if not fresh: chain = []; fresh = True
        chain.append((m.group(1).split("/")[-1], int(m.group(2)))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
    if m:
        fresh = False
        outer = [c for c in chain if c[0] == fname]
        line = outer[-1][1] if outer else -1
        key = "other"
        for nm, lo, hi in ranges:
            if lo <= line <= hi: key = nm; break
        cnt[key] += 1; tot += 1
print(f"{tot} instructions = {tot*16/1024:.1f} KB")
for k, v in cnt.most_common(): print(f"  {k:12s} {v:6d} {v*16/1024:6.1f} KB")
