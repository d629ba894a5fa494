"""Print the SASS of one kernel whose (innermost) source line lies in [lo, hi] of `file`.
usage: sass_lines.py <kernel substring> <file> <lo> <hi> [max]   (reads /tmp/w/dis.txt from tools/sass_stats.sh)"""
import re, sys
k, f, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
mx = int(sys.argv[5]) if len(sys.argv) > 5 else 400
inside = False; cur = None; n = 0
for ln in open('/tmp/w/dis.txt'):
    if ln.startswith("//---------------------"):
        inside = (k in ln) and ".text." in ln
        continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
    if m and cur and cur[0] == f and lo <= cur[1] <= hi:
        print(m.group(1), cur[1], m.group(2).strip()); n += 1
        if n >= mx: break
    elif re.match(r"\s*\.L_x_\d+:", ln):
        pass
