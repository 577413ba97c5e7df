#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tests/tools/dataflow_scale.py 5000 30000 > gpurun_out/r2c6_scale.log 2>&1
BEPUCUDA_BLOCKS_PER_SM=1 timeout 600 python tests/tools/dataflow_scale.py 30000 >> gpurun_out/r2c6_scale.log 2>&1
BEPUCUDA_DATAFLOW_NOGRAPH=1 timeout 600 python tests/tools/dataflow_scale.py 30000 >> gpurun_out/r2c6_scale.log 2>&1
cat gpurun_out/r2c6_scale.log
