#!/bin/bash
# Round 2, GPU call 2: dataflow v2 (notification counters) parity + timing + ncu source-level profile; drift tests; graph-mode baseline with rolled contacts.
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q -k "dataflow or all_execution_modes or randomised or edge_cases" 2>&1 | tail -25) > gpurun_out/r2c2_tests_dataflow.log 2>&1
(timeout 900 python -m pytest tests -m gpu -q -s -k "drift" 2>&1 | tail -25) > gpurun_out/r2c2_tests_drift.log 2>&1
echo "== graph (rolled default)" > gpurun_out/r2c2_bench.log
timeout 600 python bench.py --no-cpu-baseline --steps 30 --warmup 5 2>>gpurun_out/r2c2_bench.err | tail -1 > gpurun_out/r2c2_bench_graph.json
python -c "import json; d=json.load(open('gpurun_out/r2c2_bench_graph.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d.get('large_scene',{}).get('ms_per_step'))" >> gpurun_out/r2c2_bench.log 2>&1
for bps in 1 2; do for tune in "0,0" "200,0" "1000,0"; do
  echo "== dataflow v2 tune=$tune blocks_per_sm=$bps" >> gpurun_out/r2c2_bench.log
  BEPUCUDA_TUNE=$tune BEPUCUDA_BLOCKS_PER_SM=$bps timeout 300 python bench.py --mode dataflow --no-cpu-baseline --large-bodies 0 --steps 20 --warmup 3 2>>gpurun_out/r2c2_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'])" >> gpurun_out/r2c2_bench.log 2>&1
done; done
for scene in ragdolls fallback_stress; do
  echo "== $scene graph / dataflow" >> gpurun_out/r2c2_bench.log
  for mode in graph dataflow; do
    it=4; ss=1; bodies=160000; [ $scene = fallback_stress ] && bodies=50000
    timeout 300 python bench.py --scene $scene --bodies $bodies --substeps $ss --iterations $it --mode $mode --no-cpu-baseline --large-bodies 0 --steps 10 --warmup 3 2>>gpurun_out/r2c2_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$mode', d['ms_per_step'], d['value'])" >> gpurun_out/r2c2_bench.log 2>&1
  done
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:dataflow_solve --launch-skip 3 -c 1 -f -o gpurun_out/r2c2_dataflow_100k python bench.py --mode dataflow --no-cpu-baseline --large-bodies 0 --steps 2 --warmup 3 > gpurun_out/r2c2_ncu.log 2>&1
echo done
