#!/bin/bash
# A/B of kernel variants selected by environment variables (runs on the GPU box).  Usage: tools/gpu_ab.sh <tag> "VAR=val ..." ...
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/ab_$TAG.jsonl
for V in "$@"; do
  echo "== $V" | tee -a $OUT/ab_$TAG.jsonl
  env $V python bench.py --steps 10 --warmup 3 --no-cpu ${BENCH_ARGS:-} 2>$OUT/ab_$TAG.err | tee -a $OUT/ab_$TAG.jsonl | python -c "
import sys,json
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: print(l); continue
    print('value',round(j['value'],1),'e2e',round(j['e2e']['value'],1),'ms',j.get('kernel_ms_per_step'), 'roof', j['roofline'] and round(j['roofline']['frac'],3))
"
done
