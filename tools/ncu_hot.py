"""Hot spots of one kernel from `ncu --page source --csv`: SASS lines with the most stall samples, and a
coarse histogram of executed warp-instructions / samples per 64-instruction window."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
src, smp, ex = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
data = [(i, r[src].strip(), int(r[smp] or 0), int(r[ex] or 0)) for i, r in enumerate(rows[2:]) if len(r) > ex]
tot_s, tot_e = sum(d[2] for d in data), sum(d[3] for d in data)
print("total samples", tot_s, "warp-instructions", tot_e)
W = int(sys.argv[2]) if len(sys.argv) > 2 else 64
for w0 in range(0, len(data), W):
    s = sum(d[2] for d in data[w0:w0 + W]); e = sum(d[3] for d in data[w0:w0 + W])
    if s > 0.01 * tot_s or e > 0.01 * tot_e:
        print("%5d-%5d  samples %5.1f%%  executed %5.1f%%   %s" % (w0, w0 + W, 100 * s / tot_s, 100 * e / tot_e, data[w0][1][:50]))
print("top lines")
for d in sorted(data, key=lambda d: -d[2])[:40]:
    print("%5d %6.2f%% exec %9d  %s" % (d[0], 100 * d[2] / tot_s, d[3], d[1][:90]))
