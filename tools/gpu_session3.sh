#!/bin/bash
# round 2, session 3: validate the store-address change, per-kernel launch lists of the dense-mask and single-pair cases
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_farneback.py tests/test_gpu_round2.py -m gpu -x -q > $OUT/s3_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/s3_pytest.log
tail -5 $OUT/s3_pytest.log
timeout 600 tools/gpu_ab.sh s3 - -
timeout 600 tools/ncu_launches.sh s3_dense --workload c2dense --pairs 16 --steps 1
timeout 600 tools/ncu_launches.sh s3_c1 --workload c1 --steps 4
