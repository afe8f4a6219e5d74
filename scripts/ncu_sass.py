"""Print SASS of a profiled kernel with executed-instruction counts and stall samples (hot first or in order).
usage: python scripts/ncu_sass.py rep [--order] [--min-frac 0.005]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; order = "--order" in sys.argv
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
iA, iS, iSamp, iEx, iThr = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
data = []
for n, r in enumerate(rows[2:]):
    if len(r) <= iEx: continue
    try: data.append((n, r[iS].strip(), int(r[iSamp] or 0), int(r[iEx] or 0), r[iThr]))
    except ValueError: pass
tot_ex = sum(d[3] for d in data); tot_s = sum(d[2] for d in data)
print(f"total inst executed {tot_ex}  total samples {tot_s}  n_sass {len(data)}")
if order:
    for n, s, samp, ex, thr in data:
        print(f"{n:5d} {ex:12d} {100*ex/tot_ex:5.2f}% samp {100*samp/max(tot_s,1):5.2f}% thr {thr:>5s}  {s}")
else:
    for n, s, samp, ex, thr in sorted(data, key=lambda d: -d[2])[:60]:
        print(f"{n:5d} {ex:12d} {100*ex/tot_ex:5.2f}% samp {100*samp/max(tot_s,1):5.2f}% thr {thr:>5s}  {s}")
