mkdir -p gpurun_out
T=${1:-r2i}
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
tail -3 gpurun_out/${T}_tests.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/${T}_bench.json | head -c 600
