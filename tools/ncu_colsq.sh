#!/bin/bash
# ncu --set full of one column-norm GEMM launch (EPI_COLSQ) of the predict path: 32768 candidates against N=8192.
set -u
OUT=gpurun_out/ncu_r02
mkdir -p $OUT
python tools/ncu_target.py predict > $OUT/plain_predict.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain_predict.log; exit 1; }
ncu --clock-control none --set full --import-source on -k regex:dgemm_tasks -s 209 -c 1 -o $OUT/prof_predict_colsq python tools/ncu_target.py predict > $OUT/ncu8.log 2>&1
echo "rc=$?"; tail -3 $OUT/ncu8.log
f=$OUT/prof_predict_colsq.ncu-rep
python tools/ncu_summary.py raw $f > ${f%.ncu-rep}.txt 2>&1
ncu -i $f --page details --csv 2>/dev/null | grep -E "Duration|DRAM Throughput|Memory Throughput|Compute \(SM\) Throughput|Registers Per Thread|Achieved Occupancy|L2 Cache Throughput|Executed Ipc Active" >> ${f%.ncu-rep}.txt
rm -f $f
head -5 ${f%.ncu-rep}.txt
