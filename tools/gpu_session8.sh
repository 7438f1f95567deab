#!/bin/bash
# round 2, session 8: u32 matrices A/B (template variant), all workloads
OUT=gpurun_out; mkdir -p $OUT
timeout 600 tools/gpu_ab.sh s8 - mat_u32=1 - mat_u32=1
timeout 900 python bench.py --workload all --steps 20 > $OUT/s8_all.json 2> $OUT/s8_all.err; echo "all rc=$?"
python - <<'PY'
import json
j=json.load(open('gpurun_out/s8_all.json'))
print('c2', round(j['value'],1), 'e2e', round(j['e2e']['value'],1), 'ms', round(j['ms_per_step'],3), 'whole', j['whole_path_frac_of_hbm_peak'], {a:b['ms_per_step'] for a,b in j['kernel_classes'].items()})
print('cpu', j['cpu_baseline'])
for k,v in j['extra'].items():
    print(k, 'value %.1f'%v['value'], 'ms/step %.3f'%v['ms_per_step'], 'e2e %.1f'%v['e2e']['value'], 'whole', v['whole_path_frac_of_hbm_peak'], 'iter', v['roofline'] and round(v['roofline']['frac'],3))
PY
