"""wall-clock of the stock CLI on the C2 record (scan + ScanFold-Fold + writers + structure extraction).  usage: time_cli.py [workload]"""
import cProfile, io, os, pstats, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from scanfold_b200 import cli
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
seq, W, step, r, stype = bench.synth_record(name)
d = tempfile.mkdtemp(prefix="cli_")
open(os.path.join(d, "synth.fa"), "w").write(">synth_%s\n%s\n" % (name, seq))
os.chdir(d)
t0 = time.perf_counter()
pr = cProfile.Profile(); pr.enable()
cli.main(["synth.fa", "-w", str(W), "-s", str(step), "-r", str(r), "--type", stype])
pr.disable()
print("CLI wall time %.2f s for %d nt" % (time.perf_counter() - t0, len(seq)))
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumtime").print_stats(18); print(s.getvalue()[:3500])
print(sorted(os.listdir(os.path.join(d, "synth_%s" % name)))[:40])
