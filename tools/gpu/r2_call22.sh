#!/bin/bash
# Round 2, GPU call 22: ncu --set full of the f2 / f3 / f4 kernels (tools/ncu_aux_target.py).
mkdir -p gpurun_out
P=gpurun_out/r2c22
timeout 200 ncu --set full --clock-control none -k "regex:color_min|color_assign|color_init|predict_bounding|redistribute_impulses|scatter_body_motion|gather_body_motion" \
  -c 14 -f -o ${P}_ncu_f234 python tools/ncu_aux_target.py > ${P}_ncu_f234.log 2>&1
[ -f ${P}_ncu_f234.ncu-rep ] && ncu -i ${P}_ncu_f234.ncu-rep --page raw --csv > ${P}_ncu_f234.csv 2>/dev/null
rm -f ${P}_ncu_f234.ncu-rep
tail -2 ${P}_ncu_f234.log | cut -c1-200
python - <<PY
import csv
rows=list(csv.reader(open("${P}_ncu_f234.csv")))
h=rows[0]; k=h.index("Kernel Name"); d=h.index("gpu__time_duration.sum"); r=h.index("dram__bytes_read.sum"); w=h.index("dram__bytes_write.sum"); g=h.index("Grid Size")
for x in rows[2:]: print(x[k].split("(")[0][:40], x[g], x[d], rows[1][d], x[r], rows[1][r], x[w], rows[1][w])
PY
echo done
