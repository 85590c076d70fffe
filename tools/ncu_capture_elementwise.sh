#!/bin/bash
# ncu evidence for the element-wise kernels of the NLL path after the radial-factor stash (round 2, second half): launch
# list of one NLL+gradient evaluation, `--set full` of assemble_sym_kernel (with the stash stores) and of
# grad_contract_stash_kernel.  Same recipe as tools/ncu_capture.sh; the target first runs plain and must exit 0.
set -u
OUT=gpurun_out/ncu_r02b
mkdir -p $OUT
NCU="ncu --clock-control none"
run() { echo "== $*"; "$@"; echo "rc=$?"; }
python tools/ncu_target.py nll > $OUT/plain_nll.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain_nll.log; exit 1; }
cat $OUT/plain_nll.log
SKIP=$(grep -o "captured evaluation: [0-9]*" $OUT/plain_nll.log | grep -o "[0-9]*$")
TOTAL=$(grep -o "launches after: [0-9]*" $OUT/plain_nll.log | grep -o "[0-9]*$")
COUNT=$((TOTAL - SKIP))
echo "nll: skip $SKIP count $COUNT"
run $NCU --metrics gpu__time_duration.sum -s $SKIP -c $COUNT --csv --log-file $OUT/launches_nll_n16384.csv python tools/ncu_target.py nll > $OUT/ncu1.log 2>&1
run $NCU --set full --import-source on -k regex:assemble_sym -s 1 -c 1 -o $OUT/prof_assemble_sym_stash python tools/ncu_target.py nll > $OUT/ncu3.log 2>&1
run $NCU --set full --import-source on -k regex:grad_contract -s 1 -c 1 -o $OUT/prof_grad_contract_stash python tools/ncu_target.py nll > $OUT/ncu4.log 2>&1
run $NCU --set full --import-source on -k regex:trigemv -s 1 -c 1 -o $OUT/prof_trigemv python tools/ncu_target.py nll > $OUT/ncu5.log 2>&1
for f in $OUT/prof_*.ncu-rep; do
  python tools/ncu_summary.py raw $f > ${f%.ncu-rep}.txt 2>&1
  ncu -i $f --page details --csv 2>/dev/null | grep -E "Duration|DRAM Throughput|Memory Throughput|Compute \(SM\) Throughput|Registers Per Thread|Achieved Occupancy|L2 Cache Throughput|Executed Ipc Active" >> ${f%.ncu-rep}.txt
  rm -f $f
done
python tools/ncu_summary.py launches $OUT/launches_nll_n16384.csv > $OUT/launches_nll_n16384.txt
ls -la $OUT
for f in $OUT/ncu*.log; do echo "-- $f"; tail -n 3 $f; done
