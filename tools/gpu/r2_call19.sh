#!/bin/bash
# Round 2, GPU call 19: head-stage experiments. (a) per-SM type locality: the work records of a batch permuted so that the CTAs the block scheduler
# places on one SM walk one constraint type (BEPUCUDA_TUNE=0,1, host-side permutation only); (b) CTA size of the stage kernels: 32 / 128 threads
# (variants cta32 / cta128) against the shipped 64.
mkdir -p gpurun_out
P=gpurun_out/r2c19
(BEPUCUDA_TUNE=0,1,0,0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 240 -k "benchmark_scale or shape_pile or ragdolls or deterministic" 2>&1 | tail -4) > ${P}_tests_locality.log 2>&1
cat ${P}_tests_locality.log
S="timeout 300 python tests/tools/perf_sweep.py"
for tune in 0,0,0,0 0,1,0,0; do
  echo "== TUNE $tune C2" >> ${P}_ab.log;      SWEEP=graph BEPUCUDA_TUNE=$tune $S --bodies 100000 --steps 20 >> ${P}_ab.log 2>&1
  echo "== TUNE $tune C3 1x4" >> ${P}_ab.log;  SWEEP=graph BEPUCUDA_TUNE=$tune $S --scene ragdolls --bodies 160000 --substeps 1 --iterations 4 --steps 20 >> ${P}_ab.log 2>&1
  echo "== TUNE $tune 1M 4x2" >> ${P}_ab.log;  SWEEP=graph BEPUCUDA_TUNE=$tune $S --bodies 1000000 --substeps 4 --steps 10 >> ${P}_ab.log 2>&1
done
for v in cta32 cta128; do
  echo "== VARIANT $v C2" >> ${P}_ab.log;      SWEEP=graph BEPUCUDA_VARIANT=$v $S --bodies 100000 --steps 20 >> ${P}_ab.log 2>&1
  echo "== VARIANT $v C3 1x4" >> ${P}_ab.log;  SWEEP=graph BEPUCUDA_VARIANT=$v $S --scene ragdolls --bodies 160000 --substeps 1 --iterations 4 --steps 20 >> ${P}_ab.log 2>&1
done
grep -E "^==|^graph" ${P}_ab.log
echo done
