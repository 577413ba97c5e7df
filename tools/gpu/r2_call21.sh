#!/bin/bash
# Round 2, GPU call 21: ncu --set full of the kernels around the stage kernels that had no capture yet (f2 / f3 / f4 and the per-frame transfers).
mkdir -p gpurun_out
P=gpurun_out/r2c21
timeout 600 ncu --set full --import-source on --clock-control none -k "regex:color_min|color_assign|predict_bounding|redistribute_impulses|scatter_body_motion|gather_body_motion|pack_ref_rows|batched_copy|transpose_out_all|merge_bodies" \
  -c 40 -f -o ${P}_ncu_aux python bench.py --no-cpu-baseline --steps 1 --warmup 3 --config-steps 1 --large-bodies 0 > ${P}_ncu_aux.log 2>&1
[ -f ${P}_ncu_aux.ncu-rep ] && ncu -i ${P}_ncu_aux.ncu-rep --page raw --csv > ${P}_ncu_aux.csv 2>/dev/null
rm -f ${P}_ncu_aux.ncu-rep
ls -la gpurun_out | tail -5; tail -3 ${P}_ncu_aux.log | cut -c1-300
python - <<PY
import csv
rows=list(csv.reader(open("${P}_ncu_aux.csv")))
h=rows[0]; k=h.index("Kernel Name"); d=h.index("gpu__time_duration.sum"); r=h.index("dram__bytes_read.sum"); w=h.index("dram__bytes_write.sum"); g=h.index("Grid Size")
seen=set()
for x in rows[2:]:
    key=(x[k].split("(")[0], x[g])
    if key in seen: continue
    seen.add(key); print(x[k].split("(")[0][:40], x[g], x[d], rows[1][d], x[r], rows[1][r], x[w], rows[1][w])
PY
echo done
