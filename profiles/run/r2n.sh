mkdir -p gpurun_out
T=${1:-r2n}
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
tail -8 gpurun_out/${T}_tests.log
