#!/bin/bash
# round 2, session 5: small-tile iteration, CCL left-link variant
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/s5_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/s5_pytest.log
tail -8 $OUT/s5_pytest.log
BENCH_ARGS="--workload c1 --steps 300" timeout 600 tools/gpu_ab.sh s5c1 - iter_small_tiles=0
timeout 600 tools/gpu_ab.sh s5 - iter_small_tiles=0
BENCH_ARGS="--workload c2dense" timeout 600 tools/gpu_ab.sh s5dense -
BENCH_ARGS="--workload c2rot" timeout 600 tools/gpu_ab.sh s5rot -
BENCH_ARGS="--workload c3" timeout 600 tools/gpu_ab.sh s5c3 - iter_small_tiles=0
