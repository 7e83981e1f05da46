"""Development aid: per-launch table (time, DRAM bytes, GB/s) of an ncu --csv launch list."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
d, order = {}, []
for r in rows[hdr + 1:]:
    if len(r) < 15:
        continue
    k = int(r[0]); name = r[4].split('(')[0][-40:]; m = r[12]; v = float(r[14].replace(',', ''))
    if k not in d:
        d[k] = {'name': name, 'grid': r[8]}; order.append(k)
    d[k][m] = v
tot = 0.0
for k in order:
    e = d[k]; t = e.get('gpu__time_duration.sum', 0) / 1e3
    rd = e.get('dram__bytes_read.sum', 0) / 1e6; wr = e.get('dram__bytes_write.sum', 0) / 1e6
    tot += t
    print("%3d %-42s %-14s %9.1f us  rd %8.1f MB wr %8.1f MB  %6.0f GB/s" % (k, e['name'], e['grid'], t, rd, wr, (rd + wr) / t if t else 0))
print("total %.1f us over %d launches" % (tot, len(order)))
