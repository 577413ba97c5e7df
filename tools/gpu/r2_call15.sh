#!/bin/bash
# Round 2, GPU call 15: full GPU test-suite (packed reference rows, device colouring), bench (both arms), ncu --set full captures of every stage
# kernel selected by launch index (stream mode: the launch order is the stage program's; kernel base name, no template arguments).
mkdir -p gpurun_out
P=gpurun_out/r2c15
(time timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8) > ${P}_tests.log 2>&1
(time timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')") > ${P}_smoke.log 2>&1
(time timeout 900 python bench.py > ${P}_bench.json 2> ${P}_bench.err) 2> ${P}_bench_time.log
(time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > ${P}_bench_ref.json 2>> ${P}_bench.err) 2>> ${P}_bench_time.log
B="python bench.py --no-cpu-baseline --no-configs --steps 1 --warmup 3"
cap() { # name kernel mode skip extra-bench-args
  timeout 300 ncu --set full --import-source on --clock-control none -k $2 --launch-skip $4 -c 1 -f -o ${P}_ncu_$1 $B --mode $3 $5 > ${P}_ncu_$1.log 2>&1
}
# C2 in stream mode: one solve = 16 WarmStartFirst + 32 Solve + 7 x (1 Incremental + 16 WarmStart + 32 Solve) = 391 constraint_stage_kernel launches
cap solve constraint_stage_kernel stream 407
cap warmstart constraint_stage_kernel stream 440
cap warmstartfirst constraint_stage_kernel stream 391
cap incremental constraint_stage_kernel stream 439
cap solve_tail constraint_stage_kernel stream 422
cap dataflowpass dataflow_pass_kernel dataflow 20
# 1M bodies, 4 substeps: 18 batches, one solve = 18 + 36 + 3 x (1 + 18 + 36) = 219 launches
cap solve_1m constraint_stage_kernel stream 237 "--bodies 1000000 --substeps 4"
cap warmstart_1m constraint_stage_kernel stream 274 "--bodies 1000000 --substeps 4"
for n in solve warmstart warmstartfirst incremental solve_tail dataflowpass solve_1m warmstart_1m; do
  [ -f ${P}_ncu_$n.ncu-rep ] && ncu -i ${P}_ncu_$n.ncu-rep --page raw --csv > ${P}_ncu_$n.csv 2>/dev/null
done
[ -f ${P}_ncu_solve.ncu-rep ] && ncu -i ${P}_ncu_solve.ncu-rep --page source --csv > ${P}_ncu_solve_source.csv 2>/dev/null
[ -f ${P}_ncu_solve_1m.ncu-rep ] && ncu -i ${P}_ncu_solve_1m.ncu-rep --page source --csv > ${P}_ncu_solve_1m_source.csv 2>/dev/null
for n in warmstart warmstartfirst incremental solve_tail dataflowpass warmstart_1m solve_1m; do rm -f ${P}_ncu_$n.ncu-rep; done  # the merge back is capped at 64 MiB
ls -la gpurun_out | tail -40
cat ${P}_tests.log
echo done
