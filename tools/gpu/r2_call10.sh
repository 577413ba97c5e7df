#!/bin/bash
# Round 2, GPU call 10: full GPU test-suite, default bench (both arms), ncu --set full captures of every kernel on the path.
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8) > gpurun_out/r2c10_tests.log 2>&1
(time timeout 900 python bench.py > gpurun_out/r2c10_bench.json 2> gpurun_out/r2c10_bench.err) 2> gpurun_out/r2c10_bench_time.log
(time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c10_bench_ref.json 2>> gpurun_out/r2c10_bench.err) 2>> gpurun_out/r2c10_bench_time.log
B="python bench.py --no-cpu-baseline --no-configs --steps 1 --warmup 3"
cap() { # name regex mode skip
  timeout 300 ncu --set full --import-source on --clock-control none -k "regex:$2" --launch-skip $4 -c 1 -f -o gpurun_out/r2c10_ncu_$1 $B --mode $3 > gpurun_out/r2c10_ncu_$1.log 2>&1
}
cap solve 'constraint_stage_kernel<2, 1>' stream 40
cap warmstart 'constraint_stage_kernel<1, 1>' stream 20
cap warmstartfirst 'constraint_stage_kernel<0, 1>' stream 20
cap incremental 'constraint_stage_kernel<3, 1>' stream 3
cap finalpose 'final_pose_kernel' stream 2
cap transposein 'transpose_in_all' stream 0
cap splitbodies 'split_bodies' stream 0
cap ownership1 'ownership_pass1' stream 0
cap persistent 'persistent_solve_kernel' persistent 2
cap dataflowpass 'dataflow_pass_kernel<2' dataflow 20
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/r2c10_launches_stream_100k.csv $B --mode stream --steps 2 > /dev/null 2>&1
echo done
