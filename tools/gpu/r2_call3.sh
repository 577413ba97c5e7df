#!/bin/bash
# Round 2, GPU call 3: dataflow pass kernels (one cooperative launch per WarmStart/Solve pass): parity, timing, ncu.
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q -k "dataflow or all_execution_modes or randomised or edge_cases or kinematic_velocity" 2>&1 | tail -25) > gpurun_out/r2c3_tests_dataflow.log 2>&1
echo "== C2 dataflow" > gpurun_out/r2c3_bench.log
for bps in 0 1; do
  echo "-- blocks_per_sm=$bps" >> gpurun_out/r2c3_bench.log
  BEPUCUDA_BLOCKS_PER_SM=$bps timeout 300 python bench.py --mode dataflow --no-cpu-baseline --large-bodies 0 --no-configs --steps 30 --warmup 5 2>>gpurun_out/r2c3_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['gpu_launches'])" >> gpurun_out/r2c3_bench.log 2>&1
done
echo "== C2 dataflow strict" >> gpurun_out/r2c3_bench.log
timeout 300 python bench.py --mode dataflow --strict --no-cpu-baseline --large-bodies 0 --no-configs --steps 30 --warmup 5 2>>gpurun_out/r2c3_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'])" >> gpurun_out/r2c3_bench.log 2>&1
echo "== 1M dataflow (8x2)" >> gpurun_out/r2c3_bench.log
timeout 600 python bench.py --mode dataflow --bodies 1000000 --no-cpu-baseline --large-bodies 0 --no-configs --steps 5 --warmup 3 2>>gpurun_out/r2c3_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'])" >> gpurun_out/r2c3_bench.log 2>&1
for scene in ragdolls fallback_stress; do
  echo "== $scene graph / dataflow (1x4)" >> gpurun_out/r2c3_bench.log
  for mode in graph dataflow; do
    bodies=160000; [ $scene = fallback_stress ] && bodies=50000
    timeout 300 python bench.py --scene $scene --bodies $bodies --substeps 1 --iterations 4 --mode $mode --no-cpu-baseline --large-bodies 0 --no-configs --steps 10 --warmup 3 2>>gpurun_out/r2c3_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$mode', d['ms_per_step'], d['value'])" >> gpurun_out/r2c3_bench.log 2>&1
  done
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:dataflow_pass_kernel --launch-skip 100 -c 2 -f -o gpurun_out/r2c3_dataflow_pass_100k python bench.py --mode dataflow --no-cpu-baseline --large-bodies 0 --no-configs --steps 2 --warmup 3 > gpurun_out/r2c3_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 60 -c 120 --csv --log-file gpurun_out/r2c3_launches_dataflow.csv python bench.py --mode dataflow --no-cpu-baseline --large-bodies 0 --no-configs --steps 2 --warmup 3 > /dev/null 2>&1
echo done
