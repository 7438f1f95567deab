#!/bin/bash
# round 2, session 9: the whole GPU suite and the ncu evidence of the final build
OUT=gpurun_out; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -q > $OUT/s9_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/s9_pytest.log
tail -4 $OUT/s9_pytest.log
timeout 600 python bench.py > $OUT/s9_bench_default.json 2> $OUT/s9_bench_default.err; echo "bench rc=$?"
timeout 1500 bash tools/gpu_profile.sh r2b
python tools/timeline.py 64 > $OUT/r2_timeline_c2.txt 2>&1
