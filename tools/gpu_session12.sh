#!/bin/bash
# round 2, session 12: whole suite on the final build; the generic-iteration workloads
OUT=gpurun_out; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -q > $OUT/s12_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/s12_pytest.log
tail -3 $OUT/s12_pytest.log
BENCH_ARGS="--workload c2gauss" timeout 600 tools/gpu_ab.sh s12gauss -
BENCH_ARGS="--workload c2w21" timeout 600 tools/gpu_ab.sh s12w21 -
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
