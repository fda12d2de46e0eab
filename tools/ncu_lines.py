"""per-source-line instruction / shared-wavefront / stall-sample totals from an ncu report.  usage: ncu_lines.py rep [file-substr] [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ".cu"; top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur = None; hdr = None; out = []; agg = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if cur and want in cur and hdr:
        d = dict(zip(hdr, r))
        try:
            ins = int(d.get("Instructions Executed") or 0)
        except ValueError:
            continue
        wf = int(d.get("L1 Wavefronts Shared") or 0); smp = int(d.get("# Samples") or 0)
        if not d["Line No"]: continue
        key = (cur, int(d["Line No"]))
        if key not in agg: agg[key] = [0, 0, 0, r[1].strip()[:110]]
        agg[key][0] += ins; agg[key][1] += wf; agg[key][2] += smp
out = [(v[0], v[1], v[2], k[1], v[3]) for k, v in agg.items()]
ti = sum(o[0] for o in out); tw = sum(o[1] for o in out); ts = sum(o[2] for o in out)
print("total inst %.3g  smem wavefronts %.3g  samples %d" % (ti, tw, ts))
for o in sorted(out, key=lambda o: -o[2])[:top]:
    print("%5.1f%% inst %5.1f%% wf %5.1f%% smp  L%-4d %s" % (100 * o[0] / max(ti, 1), 100 * o[1] / max(tw, 1), 100 * o[2] / max(ts, 1), o[3], o[4]))
if len(sys.argv) > 4:
    # line ranges "name:lo-hi,lo-hi;name2:..." -> share of instructions / wavefronts / samples
    for grp in sys.argv[4].split(";"):
        name, rng = grp.split(":")
        a = b = c = 0
        for r in rng.split(","):
            lo, hi = map(int, r.split("-"))
            for o in out:
                if lo <= o[3] <= hi: a += o[0]; b += o[1]; c += o[2]
        print("%-10s inst %5.1f%%  wf %5.1f%%  smp %5.1f%%" % (name, 100 * a / ti, 100 * b / tw, 100 * c / ts))
