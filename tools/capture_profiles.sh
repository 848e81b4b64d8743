#!/bin/bash
# ncu evidence of one round: launch list of the bench command, then `--set full` of ONE pass (10 kernels) of the same command.
# usage (on the GPU box, from the repo root): bash tools/capture_profiles.sh r02_h
# Under ncu the kernels are serialised, so the lanes of the bench do not overlap there; numbers printed by these runs are
# never bench values.
TAG=${1:-cap}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-count --no-extras"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
SKIP=$(python - "$TAG" <<'PY'
import csv, sys
rows = [r for r in csv.reader(l for l in open(f'gpurun_out/{sys.argv[1]}_launches.csv') if l.startswith('"'))]
hdr = rows[0]
i_id, i_name = hdr.index('ID'), hdr.index('Kernel Name')
ids = sorted({int(r[i_id]) for r in rows[1:] if 'k_broadphase' in r[i_name]})
print(ids[5])        # the 6th pass: every lane's arena is warm
PY
)
echo "launch-skip $SKIP"
ncu --set full --clock-control none --import-source on --launch-skip $SKIP --launch-count 10 -f -o gpurun_out/${TAG}_full $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
ls -la gpurun_out/${TAG}_full.ncu-rep
