#!/bin/bash
# Round 2, GPU call 11: device-side contact update (resident impulses) parity + e2e; default bench.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_contact_update.py tests/test_gpu_parity.py -m gpu -q -k "resident or registered_host or box_stack_config1" 2>&1 | tail -15) > gpurun_out/r2c11_tests.log 2>&1
(time timeout 900 python bench.py --no-configs > gpurun_out/r2c11_bench.json 2> gpurun_out/r2c11_bench.err) 2> gpurun_out/r2c11_bench_time.log
cat gpurun_out/r2c11_tests.log
python -c "
import json; d=json.load(open('gpurun_out/r2c11_bench.json'))
for k in ('e2e','e2e_topology_change','e2e_resident_impulses'): print(k, {kk:d[k][kk] for kk in ('ms_per_step','h2d_bytes_per_step','d2h_bytes_per_step')})
print('value', d['ms_per_step'])"
tail -3 gpurun_out/r2c11_bench.err
