#!/bin/bash
# Round 2, GPU call 16: A/B of the computed work records (BEPUCUDA_TUNE=0,1 switches them off) and of the 64-register deep instantiation
# (variant deep16) on the 1M-body pile, C2 and C3; GPU test-suite on the current build.
mkdir -p gpurun_out
P=gpurun_out/r2c16
(time timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8) > ${P}_tests.log 2>&1
S="timeout 300 python tests/tools/perf_sweep.py"
for tune in 0,0,0,0 0,1,0,0; do
  echo "== TUNE $tune 1M 4x2" >> ${P}_ab.log;  SWEEP=graph BEPUCUDA_TUNE=$tune $S --bodies 1000000 --substeps 4 --steps 10 >> ${P}_ab.log 2>&1
  echo "== TUNE $tune C2" >> ${P}_ab.log;      SWEEP=graph BEPUCUDA_TUNE=$tune $S --bodies 100000 --steps 20 >> ${P}_ab.log 2>&1
  echo "== TUNE $tune C3" >> ${P}_ab.log;      SWEEP=graph BEPUCUDA_TUNE=$tune $S --scene ragdolls --bodies 160000 --substeps 1 --iterations 4 --steps 20 >> ${P}_ab.log 2>&1
done
echo "== deep16 1M 4x2" >> ${P}_ab.log; SWEEP=graph BEPUCUDA_VARIANT=deep16 $S --bodies 1000000 --substeps 4 --steps 10 >> ${P}_ab.log 2>&1
echo "== deep16 C3 8x2" >> ${P}_ab.log; SWEEP=graph BEPUCUDA_VARIANT=deep16 $S --scene ragdolls --bodies 160000 --substeps 8 --iterations 2 --steps 10 >> ${P}_ab.log 2>&1
echo "== shipped C3 8x2" >> ${P}_ab.log; SWEEP=graph $S --scene ragdolls --bodies 160000 --substeps 8 --iterations 2 --steps 10 >> ${P}_ab.log 2>&1
cat ${P}_tests.log; grep -E "^==|^graph" ${P}_ab.log
echo done
