#!/bin/bash
# weak-scaling record: bench.py at the GPU counts given (default 1 2 4 8) on one box, one JSON line each
OUT=${OUT:-gpurun_out/scaling.jsonl}
rm -f $OUT
for n in ${@:-1 2 4 8}; do
  if [ $n = 1 ]; then python bench.py --gpus 1 --steps 20 --warmup 3 --no-extras --no-count --no-cpu-baseline 2> gpurun_out/scale$n.err | grep '^{' >> $OUT
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 2> gpurun_out/scale$n.err | grep '^{' >> $OUT; fi
done
wc -l $OUT
