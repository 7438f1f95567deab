#!/bin/bash
# A/B of launch-shape variants (mavd_tuning fields) inside one GPU session.
# Usage: tools/gpu_ab.sh <tag> "field=value,field=value" ...     ("-" = the defaults)
# BENCH_ARGS adds bench.py arguments (e.g. --workload c1 --steps 200).
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/ab_$TAG.jsonl
for V in "$@"; do
  T=$V; [ "$V" = "-" ] && T=""
  echo "== $V" | tee -a $OUT/ab_$TAG.jsonl
  python bench.py --steps 10 --warmup 3 --no-cpu --tune "$T" ${BENCH_ARGS:-} 2>$OUT/ab_$TAG.err | tee -a $OUT/ab_$TAG.jsonl | python -c "
import sys,json
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: print(l); continue
    print('value',round(j['value'],1),'e2e',round(j['e2e']['value'],1),'ms/step',round(j['ms_per_step'],3),{k:v['ms_per_step'] for k,v in j['kernel_classes'].items()}, 'roof', j['roofline'] and round(j['roofline']['frac'],3))
"
done
