#!/bin/bash
# ncu evidence for profiles/ (one gpurun call; B200_PROFILING.md recipe).  Every target first runs plain and must exit 0.
set -u
OUT=gpurun_out/ncu_r02
mkdir -p $OUT
NCU="ncu --clock-control none"
run() { echo "== $*"; "$@"; echo "rc=$?"; }

python tools/ncu_target.py nll > $OUT/plain_nll.log 2>&1 && \
python tools/ncu_target.py kmat > $OUT/plain_kmat.log 2>&1 && \
python tools/ncu_target.py predict > $OUT/plain_predict.log 2>&1 || { echo "plain runs failed"; tail -5 $OUT/plain_*.log; exit 1; }
cat $OUT/plain_nll.log $OUT/plain_kmat.log $OUT/plain_predict.log

# 1. launch list of one NLL+gradient evaluation (the second one of the target)
SKIP=$(grep -o "captured evaluation: [0-9]*" $OUT/plain_nll.log | grep -o "[0-9]*$")
TOTAL=$(grep -o "launches after: [0-9]*" $OUT/plain_nll.log | grep -o "[0-9]*$")
COUNT=$((TOTAL - SKIP))
echo "nll: skip $SKIP count $COUNT"
run $NCU --metrics gpu__time_duration.sum -s $SKIP -c $COUNT --csv --log-file $OUT/launches_nll_n16384.csv python tools/ncu_target.py nll > $OUT/ncu1.log 2>&1
# index (among dgemm launches of the whole run) of the longest dgemm launch of the captured evaluation = K^-1 = U U^T,
# and of the first K>=512 trailing update
python - "$OUT/launches_nll_n16384.csv" "$SKIP" > $OUT/dgemm_index.txt <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
cols = rows[hdr]
ki, vi = cols.index("Kernel Name"), cols.index("Metric Value")
best, bi, n = 0.0, -1, 0
for r in rows[hdr + 1:]:
    if len(r) <= vi or "dgemm_tasks" not in r[ki]:
        continue
    v = float(r[vi].replace(",", ""))
    if v > best:
        best, bi = v, n
    n += 1
print(bi, n)
PY
cat $OUT/dgemm_index.txt
# dgemm launches before the captured evaluation = those of the warm-up evaluation = the same count n
IDX=$(cut -d' ' -f1 $OUT/dgemm_index.txt); NG=$(cut -d' ' -f2 $OUT/dgemm_index.txt)
run $NCU --set full --import-source on -k regex:dgemm_tasks -s $((NG + IDX)) -c 1 -o $OUT/prof_syrk python tools/ncu_target.py nll > $OUT/ncu2.log 2>&1
# 2. the element-wise kernels of the NLL path and the diagonal-block kernel (second evaluation: skip the first instances)
run $NCU --set full --import-source on -k regex:assemble_sym -s 1 -c 1 -o $OUT/prof_assemble_sym python tools/ncu_target.py nll > $OUT/ncu3.log 2>&1
run $NCU --set full --import-source on -k regex:grad_contract -s 1 -c 1 -o $OUT/prof_grad_contract python tools/ncu_target.py nll > $OUT/ncu4.log 2>&1
run $NCU --set full --import-source on -k regex:potrf_diagr -s 200 -c 1 -o $OUT/prof_potrf_diagr python tools/ncu_target.py nll > $OUT/ncu5.log 2>&1
# 3. dense K + dK (second call)
run $NCU --set full --import-source on -k regex:kernel_matrix_kernel -s 1 -c 1 -o $OUT/prof_kernel_matrix python tools/ncu_target.py kmat > $OUT/ncu6.log 2>&1
# 4. predict: cross kernel, column-norm GEMM, finish (the 65536-candidate chunk: skip the warm-up instances)
run $NCU --set full --import-source on -k regex:cross_mean -s 1 -c 1 -o $OUT/prof_cross_mean python tools/ncu_target.py predict > $OUT/ncu7.log 2>&1
# the column-norm GEMM launches are the last dgemm_tasks launches of the predict target (the 208 before them belong to the
# factorisation and the triangular inverse): 208 = warm-up chunk, 209 / 210 = the two halves of the 65536-candidate chunk
run $NCU --set full --import-source on -k regex:dgemm_tasks -s 209 -c 1 -o $OUT/prof_predict_colsq python tools/ncu_target.py predict > $OUT/ncu8.log 2>&1
run $NCU --metrics gpu__time_duration.sum -k "regex:cross_mean|dgemm_tasks|finish_variance" --csv --log-file $OUT/launches_predict.csv python tools/ncu_target.py predict > $OUT/ncu9.log 2>&1
# summarise on the box: the reports themselves are too large to travel back (64 MiB limit on gpurun_out)
for f in $OUT/prof_*.ncu-rep; do
  python tools/ncu_summary.py raw $f > ${f%.ncu-rep}.txt 2>&1
  ncu -i $f --page details --csv 2>/dev/null | grep -E "Duration|DRAM Throughput|Memory Throughput|Compute \(SM\) Throughput|Registers Per Thread|Achieved Occupancy|L2 Cache Throughput|Executed Ipc Active" >> ${f%.ncu-rep}.txt
  rm -f $f
done
python tools/ncu_summary.py launches $OUT/launches_nll_n16384.csv > $OUT/launches_nll_n16384.txt
python tools/ncu_summary.py launches $OUT/launches_predict.csv > $OUT/launches_predict.txt
ls -la $OUT
for f in $OUT/ncu*.log; do echo "-- $f"; tail -n 3 $f; done
