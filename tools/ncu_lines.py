"""Development aid: stall samples of an ncu report per CUDA source line.
usage: python tools/ncu_lines.py <source-page.csv (ncu -i rep --page source --csv)> <cubin> <kernel substring> [top]
Maps the SASS offsets of the report onto nvdisasm -g line info of the cubin."""
import csv
import re
import subprocess
import sys
from collections import defaultdict

src_csv, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
line_of = {}
cur, inside = None, False
for ln in dis:
    if ln.startswith("//---") and ".text." in ln:
        inside = kname in ln
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
base = None
per = defaultdict(lambda: [0, 0])
tot = 0
for r in rows[h + 1:]:
    if len(r) <= isamp or not r[ia].startswith("0x"):
        continue
    a = int(r[ia], 16)
    if base is None:
        base = a
    key = line_of.get(a - base, ("?", 0))
    s = int(float(r[isamp] or 0))
    per[key][0] += s
    per[key][1] += int(float(r[iex] or 0))
    tot += s
print("total samples", tot)
for key, (s, ex) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%6d %5.1f%%  inst %9d  %s:%d" % (s, 100.0 * s / max(tot, 1), ex, key[0], key[1]))
