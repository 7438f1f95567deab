#!/bin/bash
# Runs on the B200 box (via gpurun): the ncu evidence kept under profiles/.
# Usage: tools/gpu_profile.sh <tag> [bench args, default: the C2 headline workload]
#   1. the bench line without a profiler (the live CUDA-event numbers)
#   2. launch list (gpu__time_duration.sum) of `bench.py --device-only --steps 1`: whole steps of the hot path only
#   3. --set full of every kernel of the LAST captured step -> one row of counters per kernel
#   4. --set full --import-source of the heaviest kernels (one launch each) -> per-phase stall tables
# .ncu-rep files are exported to CSV and deleted (the pull limit is 64 MiB).
set -u
TAG=${1:-r2}; shift
ARGS="$@"
OUT=gpurun_out
mkdir -p $OUT
KREGEX='regex:pyr_|polyexp|matrices_init|iter_|foe_kernel|residual|seg_max|stats_|ccl_|records_fill|bgr2gray|derotate|pack_mask'
python bench.py --no-cpu $ARGS > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
BENCH="python bench.py --device-only --steps 1 --warmup 5 $ARGS"
$BENCH > $OUT/plain_$TAG.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" --csv --log-file $OUT/launches_$TAG.csv $BENCH > $OUT/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
NL=$(python - <<PY
import csv
rows=[r for r in csv.reader(open('$OUT/launches_$TAG.csv', errors='replace')) if len(r)>5 and r[0].isdigit()]
print(len(rows))
PY
)
echo "launches seen: $NL (6 steps)"
python tools/ncu_summary.py launches $OUT/launches_$TAG.csv 6 > $OUT/launches_$TAG.md
# every kernel of the last step with the full section set
ncu --set full --clock-control none -k "$KREGEX" -s $((NL*5/6)) -c $((NL/6)) -o $OUT/prof_all_$TAG -f $BENCH > $OUT/ncu_all_$TAG.log 2>&1
echo "ncu all rc=$?"
ncu -i $OUT/prof_all_$TAG.ncu-rep --page raw --csv > $OUT/prof_all_$TAG.csv 2>/dev/null
python tools/ncu_summary.py raw $OUT/prof_all_$TAG.csv > $OUT/prof_all_$TAG.md 2>/dev/null
rm -f $OUT/prof_all_$TAG.ncu-rep
# source-level capture of the heaviest kernels: the LAST launch of each (finest level of the last step)
for K in ${PROFILE_KERNELS:-iter_box_tma_kernel matrices_init polyexp_tma residual_kernel pyr_vsweep}; do
  CNT=$(python - <<PY
import csv
rows=[r for r in csv.reader(open('$OUT/launches_$TAG.csv', errors='replace')) if len(r)>5 and r[0].isdigit() and '$K' in r[4]]
best=max(range(len(rows)), key=lambda i: float(rows[i][14].replace(',',''))) if rows else 0
print(best)
PY
)
  ncu --set full --clock-control none --import-source on -k "regex:$K" -s $CNT -c 1 -o $OUT/prof_${K}_$TAG -f $BENCH > $OUT/ncu_${K}_$TAG.log 2>&1
  echo "ncu $K (launch $CNT) rc=$?"
  ncu -i $OUT/prof_${K}_$TAG.ncu-rep --page raw --csv > $OUT/prof_${K}_$TAG.raw.csv 2>/dev/null
  ncu -i $OUT/prof_${K}_$TAG.ncu-rep --page source --csv > $OUT/prof_${K}_$TAG.source.csv 2>/dev/null
  python tools/ncu_summary.py raw $OUT/prof_${K}_$TAG.raw.csv > $OUT/prof_${K}_$TAG.md 2>/dev/null
  python tools/ncu_summary.py source $OUT/prof_${K}_$TAG.source.csv >> $OUT/prof_${K}_$TAG.md 2>/dev/null
  rm -f $OUT/prof_${K}_$TAG.ncu-rep
done
ls -la $OUT | tail -30
