#!/bin/bash
# Round 2, GPU call 1: full GPU test-suite (new kinematic-velocity / benchmark-scale / fast-build drift tests), A/B of the staged kernel variants,
# dataflow-mode polling experiment.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r2c1_gpu.txt 2>&1
(time python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60) > gpurun_out/r2c1_tests.log 2>&1
python tests/tools/ab_variants.py run rolled local rolledlocal split > gpurun_out/r2c1_ab.log 2>&1
for tune in "0,0" "400,2000" "1000,5000"; do
  for bps in 1 2; do
    echo "== dataflow tune=$tune blocks_per_sm=$bps" >> gpurun_out/r2c1_dataflow.log
    BEPUCUDA_TUNE=$tune BEPUCUDA_BLOCKS_PER_SM=$bps python bench.py --mode dataflow --no-cpu-baseline --large-bodies 0 --steps 20 --warmup 3 2>>gpurun_out/r2c1_dataflow.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'])" >> gpurun_out/r2c1_dataflow.log 2>&1
  done
done
echo done
