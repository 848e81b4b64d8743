python -m pytest tests -x -q -m gpu > gpurun_out/r02n_pytest.log 2>&1; tail -2 gpurun_out/r02n_pytest.log
python bench.py > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err; echo bench rc=$?
JG_LANES=1 JG_N=3000000 JG_EDGES=200000 python tools/bench_sharded.py > gpurun_out/sh_l1.jsonl 2>gpurun_out/sh_l1.err
JG_LANES=4 JG_N=3000000 JG_EDGES=200000 python tools/bench_sharded.py > gpurun_out/sh_l4.jsonl 2>gpurun_out/sh_l4.err
bash tools/capture_profiles.sh r02_h
JG_B200_LIB_PATH=jogramop_framework_b200/libjogramop_b200_bounds.so python -m pytest tests -x -q -m gpu > gpurun_out/r02n_bounds.log 2>&1; tail -2 gpurun_out/r02n_bounds.log
