#!/bin/bash
# Round 2, GPU call 12: f2 tests green + bench with the resident-impulses e2e leg.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_contact_update.py -m gpu -q 2>&1 | tail -3) > gpurun_out/r2c12_tests.log 2>&1
(time timeout 900 python bench.py --no-configs > gpurun_out/r2c12_bench.json 2> gpurun_out/r2c12_bench.err) 2> gpurun_out/r2c12_bench_time.log
cat gpurun_out/r2c12_tests.log
python -c "
import json; d=json.load(open('gpurun_out/r2c12_bench.json'))
for k in ('e2e','e2e_topology_change','e2e_resident_impulses'): print(k, {kk:d[k][kk] for kk in ('ms_per_step','h2d_bytes_per_step','d2h_bytes_per_step')})
print('value', d['ms_per_step'])"
tail -3 gpurun_out/r2c12_bench.err
