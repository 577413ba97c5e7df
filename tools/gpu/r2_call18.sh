#!/bin/bash
# Round 2, GPU call 18: the tail-run cluster kernel (runs of small batches in one launch). Parity first (per-test timeout: a wrong mbarrier phase
# would spin forever), then A/B against the same binary with the fusion switched off (BEPUCUDA_TUNE=0,0,1).
mkdir -p gpurun_out
P=gpurun_out/r2c18
(time timeout 900 python -m pytest tests -m gpu -q -x --timeout 240 2>&1 | tail -15) > ${P}_tests.log 2>&1
cat ${P}_tests.log
if grep -q "passed" ${P}_tests.log && ! grep -q "failed\|error" ${P}_tests.log; then
  S="timeout 300 python tests/tools/perf_sweep.py"
  for tune in 0,0,0,0 0,0,1,0; do
    echo "== TUNE $tune C2" >> ${P}_ab.log;      SWEEP=graph BEPUCUDA_TUNE=$tune $S --bodies 100000 --steps 20 >> ${P}_ab.log 2>&1
    echo "== TUNE $tune C3 1x4" >> ${P}_ab.log;  SWEEP=graph BEPUCUDA_TUNE=$tune $S --scene ragdolls --bodies 160000 --substeps 1 --iterations 4 --steps 20 >> ${P}_ab.log 2>&1
    echo "== TUNE $tune C3 8x2" >> ${P}_ab.log;  SWEEP=graph BEPUCUDA_TUNE=$tune $S --scene ragdolls --bodies 160000 --substeps 8 --iterations 2 --steps 10 >> ${P}_ab.log 2>&1
    echo "== TUNE $tune C5" >> ${P}_ab.log;      SWEEP=graph BEPUCUDA_TUNE=$tune $S --scene fallback_stress --bodies 50000 --substeps 1 --iterations 4 --steps 10 >> ${P}_ab.log 2>&1
    echo "== TUNE $tune 1M 4x2" >> ${P}_ab.log;  SWEEP=graph BEPUCUDA_TUNE=$tune $S --bodies 1000000 --substeps 4 --steps 10 >> ${P}_ab.log 2>&1
  done
  grep -E "^==|^graph" ${P}_ab.log
fi
echo done
