#!/bin/bash
# Round 2, GPU call 13: full GPU test-suite, smoke, default bench (both arms), ncu --set full captures of every kernel on the path
# (current build: rolled contacts), launch list of the same command.
mkdir -p gpurun_out
P=gpurun_out/r2c13
(time timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -8) > ${P}_tests.log 2>&1
(time timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')") > ${P}_smoke.log 2>&1
(time timeout 900 python bench.py > ${P}_bench.json 2> ${P}_bench.err) 2> ${P}_bench_time.log
(time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > ${P}_bench_ref.json 2>> ${P}_bench.err) 2>> ${P}_bench_time.log
B="python bench.py --no-cpu-baseline --no-configs --steps 1 --warmup 3"
cap() { # name regex mode skip extra-bench-args
  timeout 300 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k "regex:$2" --launch-skip $4 -c 1 -f -o ${P}_ncu_$1 $B --mode $3 $5 > ${P}_ncu_$1.log 2>&1
}
cap solve 'constraint_stage_kernel<2, 1>' stream 40
cap warmstart 'constraint_stage_kernel<1, 1>' stream 20
cap warmstartfirst 'constraint_stage_kernel<0, 1>' stream 20
cap incremental 'constraint_stage_kernel<3, 1>' stream 3
cap dataflowpass 'dataflow_pass_kernel<2' dataflow 20
cap solve_1m 'constraint_stage_kernel<2, 12>' stream 36 "--bodies 1000000 --substeps 4"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 1000 --csv --log-file ${P}_launches_stream_100k.csv $B --mode stream --steps 2 > /dev/null 2>&1
for n in solve warmstart warmstartfirst incremental dataflowpass solve_1m; do
  [ -f ${P}_ncu_$n.ncu-rep ] && ncu -i ${P}_ncu_$n.ncu-rep --page raw --csv > ${P}_ncu_$n.csv 2>/dev/null
done
for n in warmstart warmstartfirst incremental dataflowpass solve_1m; do rm -f ${P}_ncu_$n.ncu-rep; done  # the merge back is capped at 64 MiB: keep the CSVs and one report
ls -la gpurun_out | tail -30
echo done
