#!/bin/bash
# round 2, session 1: tests, default bench, A/B of FUSE=2 and graph replay, all workloads
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > $OUT/s1_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s > $OUT/s1_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/s1_pytest.log
tail -15 $OUT/s1_pytest.log
timeout 600 python bench.py --steps 20 > $OUT/s1_bench.json 2> $OUT/s1_bench.err; echo "bench rc=$?"
cat $OUT/s1_bench.json | head -c 3000
timeout 900 tools/gpu_ab.sh s1 - use_graph=0 iter_fuse=2 iter_fuse=0
BENCH_ARGS="--workload c1 --steps 300" timeout 600 tools/gpu_ab.sh s1c1 - use_graph=0
timeout 900 python bench.py --workload all --steps 10 --no-cpu > $OUT/s1_all.json 2> $OUT/s1_all.err; echo "all rc=$?"
