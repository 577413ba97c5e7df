#!/bin/bash
# Round 2, GPU call 11: device-side contact update (resident impulses) parity + e2e; default bench.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_contact_update.py -m gpu -x -q 2>&1 | grep -E "AssertionError|passed|failed" | head -5; timeout 300 python tests/tools/resident_debug.py; timeout 300 python tests/tools/resident_debug.py --timings; true) > gpurun_out/r2c11_tests.log 2>&1

cat gpurun_out/r2c11_tests.log
true "
import json; d=json.load(open('gpurun_out/r2c11_bench.json'))
for k in ('e2e','e2e_topology_change','e2e_resident_impulses'): print(k, {kk:d[k][kk] for kk in ('ms_per_step','h2d_bytes_per_step','d2h_bytes_per_step')})
print('value', d['ms_per_step'])"
tail -3 gpurun_out/r2c11_bench.err
