"""dev tool: aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows[:10]) if 'Instructions Executed' in r][0]
hdr = rows[hi]
ci = hdr.index('Instructions Executed'); si = hdr.index('# Samples')
data = []
for r in rows[hi + 1:]:
    if len(r) <= ci or r[0] == '':
        continue
    try:
        data.append((int(r[ci]), int(r[si]), int(r[0]), r[1].strip()[:115]))
    except ValueError:
        pass
tot = sum(d[0] for d in data); tots = sum(d[1] for d in data)
print("total inst", tot, "samples", tots)
key = int(sys.argv[2]) if len(sys.argv) > 2 else 1
data.sort(key=lambda d: -d[key])
for d in data[:int(sys.argv[3]) if len(sys.argv) > 3 else 50]:
    print("%8d inst %5.1f%%  %6d samp %5.1f%%  L%-4d %s" % (d[0], 100 * d[0] / tot, d[1], 100 * d[1] / max(tots, 1), d[2], d[3]))
