#!/bin/bash
# Runs on the B200 box (via gpurun): GPU tests, the bench, the per-launch time list and ncu captures.
# Usage: tools/gpu_profile.sh <tag> [skip_tests]
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
KREGEX='regex:pyr_|polyexp|matrices_init|iter_|foe_kernel|residual|seg_max|stats_|ccl_|records_fill|frame_'
if [ "${2:-}" != "skip_tests" ]; then
  python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_$TAG.log
  tail -3 $OUT/pytest_$TAG.log
fi
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
cat $OUT/bench_$TAG.json
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu"
$BENCH > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" --csv --log-file $OUT/launches_$TAG.csv $BENCH > $OUT/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
# one step (after 3 warm-up steps) of every kernel with the bandwidth / occupancy sections
NL=$(python - <<PY
import csv,sys
rows=[r for r in csv.reader(open('$OUT/launches_$TAG.csv')) if len(r)>5 and r[0].isdigit()]
print(len(rows))
PY
)
echo "launches seen: $NL"
ncu --set full --clock-control none -k "$KREGEX" -s $((NL*3/8)) -c $((NL/8)) -o $OUT/prof_all_$TAG -f $BENCH > $OUT/ncu_all_$TAG.log 2>&1
echo "ncu all rc=$?"
ncu -i $OUT/prof_all_$TAG.ncu-rep --page raw --csv > $OUT/prof_all_$TAG.csv 2>/dev/null
rm -f $OUT/prof_all_$TAG.ncu-rep    # the CSV travels back, the report is too large (64 MiB pull limit)
# source-level capture of the heaviest kernels (one launch each, after the warm-up steps)
for KS in ${PROFILE_KERNELS:-iter_box_tma:69 residual_kernel:4 pyr_hpass:3 polyexp:23 matrices_init:23}; do
  K=${KS%%:*}; S=${KS##*:}
  ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -o $OUT/prof_${K}_$TAG -f $BENCH > $OUT/ncu_${K}_$TAG.log 2>&1
  echo "ncu $K rc=$?"
done
ls -la $OUT
