#!/bin/bash
# round 2, session 7: border-tile fix-up, u32 matrices A/B
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_farneback.py tests/test_gpu_round2.py tests/test_gpu_process.py -m gpu -q > $OUT/s7_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/s7_pytest.log
tail -4 $OUT/s7_pytest.log
timeout 600 tools/gpu_ab.sh s7 - mat_u32=1 -
BENCH_ARGS="--workload c3" timeout 600 tools/gpu_ab.sh s7c3 - mat_u32=1
BENCH_ARGS="--workload c1 --steps 300" timeout 600 tools/gpu_ab.sh s7c1 -
