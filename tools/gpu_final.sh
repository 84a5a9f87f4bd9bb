#!/bin/bash
# Round-end style measurement on the GPU box: the product arm and the reference arm exactly as the driver launches them.
#   gpurun [--gpus N] --timeout 1200 -- 'bash tools/gpu_final.sh r02 N'
TAG=${1:-rXX}
N=${2:-1}
OUT=gpurun_out
mkdir -p $OUT
if [ "$N" = "1" ]; then
  S=$(date +%s); timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
  echo "ours n=1 rc=$? $(( $(date +%s) - S ))s"
  S=$(date +%s); timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/${TAG}_reference_n1.json 2> $OUT/${TAG}_reference_n1.err
  echo "reference n=1 rc=$? $(( $(date +%s) - S ))s"
else
  S=$(date +%s); timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
      bench.py --gpus $N --steps 20 --warmup 5 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err
  echo "ours n=$N rc=$? $(( $(date +%s) - S ))s"
fi
python tools/show_bench.py $OUT/${TAG}_bench_n$N.json
