"""Per-source-line summary of an ncu report (manual tool): python tests/ncu_lines.py rep.ncu-rep [top [kernel-regex]]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kern = ["-k", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else []
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + kern, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
hdr = rows[hi]
iI = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); iT = hdr.index("Thread Instructions Executed")
cur = None; agg = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1]; continue
    if r[0] in ("Function Name", "Line No") or r[0] == "": continue
    try: agg.append((cur.split('/')[-1], int(r[0]), r[1].strip()[:100], int(r[iI] or 0), int(r[iS] or 0), int(r[iT] or 0)))
    except ValueError: pass
tot = sum(a[3] for a in agg); ts = sum(a[4] for a in agg)
print("warp instructions", tot, "samples", ts)
for a in sorted(agg, key=lambda a: -a[3])[:top]:
    print("%-16s %5d inst %5.1f%% samp %5.1f%% thr %4.1f  %s" % (a[0], a[1], 100 * a[3] / tot, 100 * a[4] / max(ts, 1), a[5] / max(a[3], 1), a[2]))
