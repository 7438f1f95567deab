#!/bin/bash
# round 2, session 4: TMA polyexp + CCL pre-link: full tests, A/B, dense / rotation workloads
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/s4_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/s4_pytest.log
tail -8 $OUT/s4_pytest.log
timeout 600 tools/gpu_ab.sh s4 - polyexp_tma=0
BENCH_ARGS="--workload c2dense" timeout 600 tools/gpu_ab.sh s4dense -
BENCH_ARGS="--workload c2rot" timeout 600 tools/gpu_ab.sh s4rot -
timeout 600 tools/ncu_launches.sh s4_dense --workload c2dense --pairs 16 --steps 1
