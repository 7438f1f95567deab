#!/bin/bash
# Runs on the B200 box (via gpurun): GPU tests, the bench, the per-launch time list and ncu captures.
# Usage: tools/gpu_profile.sh <tag> [skip_tests] [skip_ncu]
# Everything lands in gpurun_out/ as text (CSV / logs); the .ncu-rep files are exported and deleted
# because the pull limit is 64 MiB.
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
KREGEX='regex:pyr_|polyexp|matrices_init|iter_|foe_kernel|residual|seg_max|stats_|ccl_|records_fill|frame_|bgr2gray|derotate'
if [ "${2:-}" != "skip_tests" ]; then
  python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_$TAG.log
  tail -3 $OUT/pytest_$TAG.log
fi
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
cat $OUT/bench_$TAG.json
if [ "${3:-}" = "skip_ncu" ]; then exit 0; fi
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu"
$BENCH > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" --csv --log-file $OUT/launches_$TAG.csv $BENCH > $OUT/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
NL=$(python - <<PY
import csv
rows=[r for r in csv.reader(open('$OUT/launches_$TAG.csv', errors='replace')) if len(r)>5 and r[0].isdigit()]
print(len(rows))
PY
)
echo "launches seen: $NL"
# one device step (the 4th of 8: 4 device steps then 4 host-path steps) of every kernel with the full section set
ncu --set full --clock-control none -k "$KREGEX" -s $((NL*3/8)) -c $((NL/8)) -o $OUT/prof_all_$TAG -f $BENCH > $OUT/ncu_all_$TAG.log 2>&1
echo "ncu all rc=$?"
ncu -i $OUT/prof_all_$TAG.ncu-rep --page raw --csv > $OUT/prof_all_$TAG.csv 2>/dev/null
rm -f $OUT/prof_all_$TAG.ncu-rep
# source-level capture of the heaviest kernels (one launch each, after the warm-up steps): <name regex>:<launches to skip>
for KS in ${PROFILE_KERNELS:-iter_box_tma:69 matrices_init:23 pyr_vfirst:3 polyexp:22 residual_kernel:4}; do
  K=${KS%%:*}; S=${KS##*:}
  N=$(echo $K | tr -cd 'a-z_')
  ncu --set full --clock-control none --import-source on -k "regex:$K" -s $S -c 1 -o $OUT/prof_${N}_$TAG -f $BENCH > $OUT/ncu_${N}_$TAG.log 2>&1
  echo "ncu $N rc=$?"
  ncu -i $OUT/prof_${N}_$TAG.ncu-rep --page raw --csv > $OUT/prof_${N}_$TAG.raw.csv 2>/dev/null
  ncu -i $OUT/prof_${N}_$TAG.ncu-rep --page source --csv > $OUT/prof_${N}_$TAG.source.csv 2>/dev/null
  rm -f $OUT/prof_${N}_$TAG.ncu-rep
done
ls -la $OUT
