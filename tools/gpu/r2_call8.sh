#!/bin/bash
# Round 2, GPU call 5: dataflow pass kernels at scale (co-residency fix), timing, ncu.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -k "dataflow or all_execution_modes or randomised or edge_cases or kinematic_velocity" 2>&1 | tail -5) > gpurun_out/r2c8_tests.log 2>&1
timeout 900 python tests/tools/dataflow_scale.py 30000 100000 > gpurun_out/r2c8_scale.log 2>&1
echo "== C2 dataflow" > gpurun_out/r2c8_bench.log
for bps in 0 1; do
  echo "-- blocks_per_sm=$bps" >> gpurun_out/r2c8_bench.log
  BEPUCUDA_BLOCKS_PER_SM=$bps timeout 300 python bench.py --mode dataflow --no-cpu-baseline --large-bodies 0 --no-configs --steps 30 --warmup 5 2>>gpurun_out/r2c8_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d.get('e2e_topology_change',{}).get('ms_per_step'))" >> gpurun_out/r2c8_bench.log 2>&1
done
echo "== C2 dataflow strict" >> gpurun_out/r2c8_bench.log
timeout 300 python bench.py --mode dataflow --strict --no-cpu-baseline --large-bodies 0 --no-configs --steps 30 --warmup 5 2>>gpurun_out/r2c8_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'])" >> gpurun_out/r2c8_bench.log 2>&1
echo "== 1M dataflow (8x2)" >> gpurun_out/r2c8_bench.log
timeout 600 python bench.py --mode dataflow --bodies 1000000 --no-cpu-baseline --large-bodies 0 --no-configs --steps 5 --warmup 3 2>>gpurun_out/r2c8_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'])" >> gpurun_out/r2c8_bench.log 2>&1
for scene in ragdolls fallback_stress; do
  echo "== $scene dataflow (1x4)" >> gpurun_out/r2c8_bench.log
  bodies=160000; [ $scene = fallback_stress ] && bodies=50000
  timeout 300 python bench.py --scene $scene --bodies $bodies --substeps 1 --iterations 4 --mode dataflow --no-cpu-baseline --large-bodies 0 --no-configs --steps 10 --warmup 3 2>>gpurun_out/r2c8_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('dataflow', d['ms_per_step'], d['value'])" >> gpurun_out/r2c8_bench.log 2>&1
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:dataflow_pass_kernel --launch-skip 100 -c 2 -f -o gpurun_out/r2c8_dataflow_pass_100k python bench.py --mode dataflow --no-cpu-baseline --large-bodies 0 --no-configs --steps 2 --warmup 3 > gpurun_out/r2c8_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 60 -c 120 --csv --log-file gpurun_out/r2c8_launches_dataflow.csv python bench.py --mode dataflow --no-cpu-baseline --large-bodies 0 --no-configs --steps 2 --warmup 3 > /dev/null 2>&1
echo done
