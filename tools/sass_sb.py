"""Decode scoreboard (write-barrier / wait-mask) assignment of the 128-bit loads in a kernel.
usage: sass_sb.py <so-or-cubin> <kernel substring> [opcode regex]"""
import re, subprocess, sys
f, k = sys.argv[1], sys.argv[2]
pat = re.compile(sys.argv[3] if len(sys.argv) > 3 else r'LDG\.E\.128')
lines = subprocess.run(["cuobjdump", "-sass", f], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(lines) if 'Function :' in l and k in l)
ins = []
i = start + 1
while i < len(lines) and 'Function :' not in lines[i]:
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/', lines[i])
    if m and i + 1 < len(lines):
        m2 = re.match(r'\s+/\* (0x[0-9a-f]+) \*/', lines[i + 1])
        ctrl = int(m2.group(1), 16) >> 41
        ins.append((int(m.group(1), 16), m.group(2), ctrl & 0xf, (ctrl >> 5) & 7, (ctrl >> 8) & 7, (ctrl >> 11) & 0x3f))
        i += 2
    else:
        i += 1
k0 = next(n for n, t in enumerate(ins) if pat.search(t[1]))
k1 = max(n for n, t in enumerate(ins) if pat.search(t[1]))
for t in ins[k0:k1 + 40]:
    if pat.search(t[1]) or (t[5] & 0x3c) or 'BRA' in t[1]:
        print(hex(t[0]), t[1][:58].ljust(58), 'wr', t[3] if t[3] != 7 else '-', 'wait', format(t[5], '06b'))
