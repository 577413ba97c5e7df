#!/bin/bash
# Round 2, GPU call 9: dataflow pass kernels with strong gpu-scope record accesses, no fence (A/B with the fence knob).
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -k "dataflow or all_execution_modes or randomised or edge_cases or kinematic_velocity" 2>&1 | tail -5) > gpurun_out/r2c9_tests.log 2>&1
timeout 900 python tests/tools/dataflow_scale.py 30000 100000 > gpurun_out/r2c9_scale.log 2>&1
echo "== C2 dataflow (no fence)" > gpurun_out/r2c9_bench.log
for bps in 0 1; do
  echo "-- blocks_per_sm=$bps" >> gpurun_out/r2c9_bench.log
  BEPUCUDA_BLOCKS_PER_SM=$bps timeout 300 python bench.py --mode dataflow --no-cpu-baseline --large-bodies 0 --no-configs --steps 30 --warmup 5 2>>gpurun_out/r2c9_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d.get('e2e_topology_change',{}).get('ms_per_step'))" >> gpurun_out/r2c9_bench.log 2>&1
done
echo "== C2 dataflow with release fence" >> gpurun_out/r2c9_bench.log
BEPUCUDA_TUNE=0,0,1,0 timeout 300 python bench.py --mode dataflow --no-cpu-baseline --large-bodies 0 --no-configs --steps 30 --warmup 5 2>>gpurun_out/r2c9_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'])" >> gpurun_out/r2c9_bench.log 2>&1
echo "== C2 dataflow strict" >> gpurun_out/r2c9_bench.log
timeout 300 python bench.py --mode dataflow --strict --no-cpu-baseline --large-bodies 0 --no-configs --steps 30 --warmup 5 2>>gpurun_out/r2c9_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'])" >> gpurun_out/r2c9_bench.log 2>&1
echo "== 1M dataflow (8x2)" >> gpurun_out/r2c9_bench.log
timeout 600 python bench.py --mode dataflow --bodies 1000000 --no-cpu-baseline --large-bodies 0 --no-configs --steps 5 --warmup 3 2>>gpurun_out/r2c9_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'])" >> gpurun_out/r2c9_bench.log 2>&1
for scene in ragdolls fallback_stress; do
  echo "== $scene dataflow (1x4)" >> gpurun_out/r2c9_bench.log
  bodies=160000; [ $scene = fallback_stress ] && bodies=50000
  timeout 300 python bench.py --scene $scene --bodies $bodies --substeps 1 --iterations 4 --mode dataflow --no-cpu-baseline --large-bodies 0 --no-configs --steps 10 --warmup 3 2>>gpurun_out/r2c9_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('dataflow', d['ms_per_step'], d['value'])" >> gpurun_out/r2c9_bench.log 2>&1
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:dataflow_pass_kernel --launch-skip 100 -c 2 -f -o gpurun_out/r2c9_dataflow_pass_100k python bench.py --mode dataflow --no-cpu-baseline --large-bodies 0 --no-configs --steps 2 --warmup 3 > gpurun_out/r2c9_ncu.log 2>&1
echo done
