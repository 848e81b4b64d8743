rm -f gpurun_out/r02k_scaling.jsonl
python bench.py --gpus 1 --steps 20 --warmup 3 --no-extras --no-count --no-cpu-baseline >> gpurun_out/r02k_scaling.jsonl 2> gpurun_out/r02k_scale1.err
for n in 2 4 8; do python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 >> gpurun_out/r02k_scaling.jsonl 2> gpurun_out/r02k_scale$n.err; done
wc -l gpurun_out/r02k_scaling.jsonl
