#!/bin/bash
# round 2, session 6: the whole GPU suite (no -x), then the ncu evidence of the C2 headline
OUT=gpurun_out; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -q > $OUT/s6_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/s6_pytest.log
tail -12 $OUT/s6_pytest.log
timeout 1500 bash tools/gpu_profile.sh r2a
