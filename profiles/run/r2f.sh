mkdir -p gpurun_out
T=${1:-r2f}
timeout 900 python -m pytest tests/test_gpu_bloom.py -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
tail -3 gpurun_out/${T}_tests.log
timeout 600 python bench.py --stages construct --no-cpu-baseline > gpurun_out/${T}_construct.json 2> gpurun_out/${T}_construct.err; echo "bench rc=$?"
python - <<PY
import json
f="gpurun_out/${T}_construct.json"
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    st=d["stages"]["construct"]
    print(f, d["value"], st["kernel_ms_per_step"], st["e2e"]["value"], st["result"])
except Exception as e:
    print(f, "ERR", e)
PY
python profiles/run/construct_once.py 1000000 2 > gpurun_out/${T}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"ft_(hash|append|resolve)" -s 3 -c 3 -o gpurun_out/${T}_ft python profiles/run/construct_once.py 1000000 2 > gpurun_out/${T}_ncu.log 2>&1
echo ncu rc=$?
