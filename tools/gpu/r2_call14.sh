#!/bin/bash
# Round 2, GPU call 14: ncu --set full captures of every stage kernel (kernel names as ncu demangles them: "<(int)2, (int)1>"),
# launch list of the same command, execution-mode sweep on C2 / C3 / 1M (development numbers for profiles/r02_summary.md).
mkdir -p gpurun_out
P=gpurun_out/r2c14
B="python bench.py --no-cpu-baseline --no-configs --steps 1 --warmup 3"
cap() { # name regex mode skip extra-bench-args
  timeout 300 ncu --set full --import-source on --clock-control none -k "regex:$2" --launch-skip $4 -c 1 -f -o ${P}_ncu_$1 $B --mode $3 $5 > ${P}_ncu_$1.log 2>&1
}
cap solve 'constraint_stage_kernel<\(int\)2, \(int\)1>' stream 40
cap warmstart 'constraint_stage_kernel<\(int\)1, \(int\)1>' stream 20
cap warmstartfirst 'constraint_stage_kernel<\(int\)0, \(int\)1>' stream 20
cap incremental 'constraint_stage_kernel<\(int\)3, \(int\)1>' stream 3
cap dataflowpass 'dataflow_pass_kernel<\(int\)2' dataflow 20
cap solve_1m 'constraint_stage_kernel<\(int\)2, \(int\)12>' stream 36 "--bodies 1000000 --substeps 4"
for n in solve warmstart warmstartfirst incremental dataflowpass solve_1m; do
  [ -f ${P}_ncu_$n.ncu-rep ] && ncu -i ${P}_ncu_$n.ncu-rep --page raw --csv > ${P}_ncu_$n.csv 2>/dev/null
done
[ -f ${P}_ncu_solve.ncu-rep ] && ncu -i ${P}_ncu_solve.ncu-rep --page source --csv > ${P}_ncu_solve_source.csv 2>/dev/null
for n in warmstart warmstartfirst incremental dataflowpass solve_1m; do rm -f ${P}_ncu_$n.ncu-rep; done  # the merge back is capped at 64 MiB
(SWEEP=full timeout 300 python tests/tools/perf_sweep.py --bodies 100000 --steps 20) > ${P}_sweep_c2.log 2>&1
(SWEEP=dataflow timeout 300 python tests/tools/perf_sweep.py --bodies 100000 --steps 20) > ${P}_sweep_c2_dataflow.log 2>&1
(SWEEP=full timeout 300 python tests/tools/perf_sweep.py --scene ragdolls --bodies 160000 --substeps 1 --iterations 4 --steps 20) > ${P}_sweep_c3.log 2>&1
(SWEEP=graph timeout 300 python tests/tools/perf_sweep.py --bodies 1000000 --substeps 4 --steps 10) > ${P}_sweep_1m.log 2>&1
ls -la gpurun_out | tail -30
tail -8 ${P}_sweep_c2.log ${P}_sweep_c2_dataflow.log ${P}_sweep_c3.log ${P}_sweep_1m.log
echo done
