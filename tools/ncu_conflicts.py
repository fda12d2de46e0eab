"""shared-memory bank conflicts per source line from an ncu report.  usage: ncu_conflicts.py rep [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 20
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None; agg = {}
for r in rows:
    if not r: continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0].isdigit():
        d = dict(zip(hdr, r))
        try:
            wf = int(d.get("L1 Wavefronts Shared") or 0); ex = int(d.get("L1 Wavefronts Shared Excessive") or 0); ins = int(d.get("Instructions Executed") or 0)
        except ValueError:
            continue
        a = agg.setdefault(int(r[0]), [0, 0, 0, r[1].strip()[:100]]); a[0] += ex; a[1] += wf; a[2] += ins
tw = sum(a[1] for a in agg.values()); te = sum(a[0] for a in agg.values()); ti = sum(a[2] for a in agg.values())
print("total wavefronts %.4g  excessive %.4g  instructions %.4g" % (tw, te, ti))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("wf %5.1f%%  excess %5.1f%% (%.2fx)  inst %4.1f%%  L%d %s" % (100 * a[1] / tw, 100 * a[0] / max(te, 1), a[1] / max(a[1] - a[0], 1), 100 * a[2] / ti, k, a[3]))
