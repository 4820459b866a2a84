"""Aggregate one stall reason of an ncu source page by CUDA source line: python tools/ncu_stall_lines.py rep stall_no_inst [top [launch]]; NCU_FILTER="-k regex:name -s 2 -c 1" selects launches of the report"""
import csv, subprocess, sys
rep, col = sys.argv[1], sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 30; want = int(sys.argv[4]) if len(sys.argv) > 4 else 1
import os, shlex
out = subprocess.run(["ncu", "-i", rep] + shlex.split(os.environ.get("NCU_FILTER", "")) + ["--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = None; agg = {}; launch = 0
for r in rows:
    if len(r) > 8 and r[0] == "Line No":
        launch += 1   # (one header per source file of a launch: with NCU_FILTER every section is taken)
        h = r if (col in r and "# Samples" in r) else None   # (launches too short to be sampled have no such columns)
        if h: ci = h.index(col); sm = h.index("# Samples")
        continue
    if h is None or len(r) != len(h) or r[2] != "-" or (launch != want and "NCU_FILTER" not in os.environ):
        continue
    try:
        key = (int(r[0]), r[1]); a = agg.setdefault(key, [0.0, 0.0]); a[0] += float(r[ci]); a[1] += float(r[sm])
    except ValueError:
        pass
tot = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print("%s: %.0f of %.0f samples (%.1f%%), launch %d of the report" % (col, tot, ts, 100 * tot / max(ts, 1), want))
for (ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5d  %5.1f%% of %s, %5.1f%% of the line's samples | %s" % (ln, 100 * a[0] / max(tot, 1), col, 100 * a[0] / max(a[1], 1), src.strip()[:100]))
