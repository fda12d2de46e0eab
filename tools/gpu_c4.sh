#!/bin/bash
# C4 (BASELINE.json configs[3]): 1 Mb record, 120-nt window, step 1, 100 shuffles, through the CLI on N GPUs (torchrun),
# wall time and peak device memory; then the N-GPU bench line.   usage (under gpurun --gpus N): bash tools/gpu_c4.sh N tag
N=${1:-8}; TAG=${2:-r02c4}; OUT=$PWD/gpurun_out/$TAG; mkdir -p $OUT; REPO=$PWD
work=$(mktemp -d)
python - <<PY
import sys; sys.path.insert(0, "$REPO/tools")
from bench_configs import synth
open("$work/c4.fa", "w").write(">synth_C4\n" + synth(1000000, 1004) + "\n")
PY
( while true; do nvidia-smi --query-gpu=index,memory.used --format=csv,noheader,nounits; sleep 1; done ) > $OUT/mem_samples.csv 2>/dev/null &
SAMPLER=$!
cd $work
T0=$(date +%s.%N)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 \
    $REPO/ScanFold.py c4.fa > $OUT/cli_c4.log 2> $OUT/cli_c4.err
echo "cli rc=$? wall $(python -c "import time; print(round(time.time() - $T0, 1))") s" | tee -a $OUT/cli_c4.log
kill $SAMPLER
cd $REPO
grep -E "Elapsed|Total runtime|complete" $OUT/cli_c4.log | tail -5
ls -la $work/synth_C4 | head -40 > $OUT/c4_files.txt; du -sh $work/synth_C4 >> $OUT/c4_files.txt
python - <<PY
import collections
peak = collections.defaultdict(int)
for ln in open("$OUT/mem_samples.csv"):
    f = ln.split(",")
    if len(f) == 2: peak[int(f[0])] = max(peak[int(f[0])], int(f[1]))
print("peak device memory (MiB) per GPU:", dict(peak))
PY
head -c 600 $work/synth_C4/*.out | head -5 | cut -c1-200
if [ "$3" = "bench" ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29622 \
    bench.py --gpus $N --steps 3 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
grep -v "^NCCL" $OUT/bench_n$N.json | head -c 1500
fi
