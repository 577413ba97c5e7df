#!/bin/bash
# Round 2, GPU call 20: full GPU test-suite (incl. PredictBoundingBoxes and colouring), smoke, bench both arms on the current build.
mkdir -p gpurun_out
P=gpurun_out/r2c20
(time timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8) > ${P}_tests.log 2>&1
(time timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')") > ${P}_smoke.log 2>&1
(time timeout 900 python bench.py > ${P}_bench.json 2> ${P}_bench.err) 2> ${P}_bench_time.log
(time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > ${P}_bench_ref.json 2>> ${P}_bench.err) 2>> ${P}_bench_time.log
cat ${P}_tests.log ${P}_smoke.log; tail -3 ${P}_bench.err; head -c 600 ${P}_bench.json
echo done
