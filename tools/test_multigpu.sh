#!/bin/bash
# N-GPU check (run under `gpurun --gpus N`): CLI outputs of an N-rank torchrun run must equal the golden files,
# then the N-rank bench line.   usage: bash tools/test_multigpu.sh N tag
N=${1:-2}; TAG=${2:-mgpu}
OUT=$PWD/gpurun_out/$TAG; mkdir -p $OUT
REPO=$PWD
for case in mono_w60_gc hc_w40 mono_w40 c1_w120_r100 q10_n260_w120 by_ed_w40; do
  work=$(mktemp -d); cp tests/golden/$case/input.fa tests/golden/$case/trace.npz $work/
  [ -f tests/golden/$case/constraints.dbn ] && cp tests/golden/$case/constraints.dbn $work/
  args=$(python -c "
import json,sys; a=json.load(open('tests/golden/$case/case.json'))['args']
print(' '.join(x if x!='constraints.dbn' else '$work/constraints.dbn' for x in a))")
  (cd $work && python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
      $REPO/ScanFold.py input.fa $args --parity_shuffles trace.npz > $OUT/cli_$case.log 2>&1)
  rec=$(python -c "import json; print(json.load(open('tests/golden/$case/case.json'))['record'])")
  nbad=0
  for f in tests/golden/$case/expected/*; do
    b=$(basename $f)
    case $b in *motif*|ExtractedStructures.gff3) ;; esac
    cmp -s $f $work/$rec/$b || { echo "DIFF $case $b"; nbad=$((nbad+1)); }
  done
  echo "case $case on $N GPUs: $nbad differing files" | tee -a $OUT/summary.txt
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 \
    bench.py --gpus $N --steps 3 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
tail -3 $OUT/bench_n$N.err; cat $OUT/bench_n$N.json | head -c 1500
