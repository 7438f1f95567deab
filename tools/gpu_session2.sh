#!/bin/bash
# round 2, session 2: full GPU tests (new CCL, vis, flo), bench c2 + all workloads
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q -s > $OUT/s2_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/s2_pytest.log
tail -15 $OUT/s2_pytest.log
timeout 900 python bench.py --workload all --steps 10 --no-cpu > $OUT/s2_all.json 2> $OUT/s2_all.err; echo "all rc=$?"
python - <<'PY'
import json
j=json.load(open('gpurun_out/s2_all.json'))
print('c2', round(j['value'],1), 'e2e', round(j['e2e']['value'],1), {a:b['ms_per_step'] for a,b in j['kernel_classes'].items()})
for k,v in j['extra'].items():
    print(k, 'value %.1f'%v['value'], 'ms/step %.3f'%v['ms_per_step'], 'e2e %.1f'%v['e2e']['value'], {a:b['ms_per_step'] for a,b in v['kernel_classes'].items()})
PY
