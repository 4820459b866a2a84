"""Aggregate an ncu source page (cuda,sass view) by CUDA source line: samples, instructions, lane efficiency."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = None; agg = {}
for r in rows:
    if len(r) > 8 and r[0] == "Line No":
        h = r; ie = h.index("Instructions Executed"); te = h.index("Thread Instructions Executed"); sm = h.index("# Samples"); continue
    if h is None or len(r) != len(h) or r[2] != "-":   # keep the per-line aggregate rows (Address == '-')
        continue
    try:
        key = (int(r[0]), r[1])
        a = agg.setdefault(key, [0.0, 0.0, 0.0])
        a[0] += float(r[sm]); a[1] += float(r[ie]); a[2] += float(r[te])
    except ValueError:
        pass
ts = sum(a[0] for a in agg.values()); ti = sum(a[1] for a in agg.values()); tt = sum(a[2] for a in agg.values())
print("total samples %.0f warp-inst %.3g thread-inst %.3g  lanes/inst %.1f" % (ts, ti, tt, tt / ti))
for (ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5d  %5.1f%% smp  %5.1f%% inst  lanes %4.1f | %s" % (ln, 100 * a[0] / ts, 100 * a[1] / ti, a[2] / max(a[1], 1), src.strip()[:105]))
