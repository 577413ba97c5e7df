#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2c7_scale.log
echo "== plain" > $L
timeout 600 python tests/tools/dataflow_scale.py 20000 24000 >> $L 2>&1
echo "== release fence between record and notification" >> $L
BEPUCUDA_TUNE=0,0,1,0 timeout 600 python tests/tools/dataflow_scale.py 24000 30000 >> $L 2>&1
echo "== both fences" >> $L
BEPUCUDA_TUNE=0,0,1,1 timeout 600 python tests/tools/dataflow_scale.py 24000 30000 >> $L 2>&1
echo "== memcheck 24000" >> $L
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python tests/tools/dataflow_scale.py 24000 2>&1 | tail -30 >> $L
cat $L
