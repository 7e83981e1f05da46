"""Summarise an `ncu --page source --csv` dump: hottest SASS instructions by stall samples.
usage: python tools/ncu_top.py src.csv [top_n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {k: i for i, k in enumerate(hdr)}
body = rows[2:]
S = ix["# Samples"]
tot = sum(int(r[S]) for r in body)
stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
agg = {k: sum(int(r[ix[k]]) for r in body) for k in stalls}
print("total samples", tot, " instructions", len(body))
print("by reason:", ", ".join("%s=%.1f%%" % (k[6:], 100.0 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
order = sorted(range(len(body)), key=lambda i: -int(body[i][S]))[:top]
for i in sorted(order):
    r = body[i]
    why = sorted(((int(r[ix[k]]), k[6:]) for k in stalls), reverse=True)[:2]
    print("%5d %5.1f%% exec=%-8s %-70s %s" % (i, 100.0 * int(r[S]) / tot, r[ix["Instructions Executed"]], r[ix["Source"]].strip()[:70],
                                    " ".join("%s:%d" % (n, c) for c, n in why if c)))
