mkdir -p gpurun_out
T=${1:-r2k}
timeout 1200 python -m pytest tests/test_gpu_packed.py tests/test_gpu_bloom.py tests/test_gpu_fullsize.py tests/test_gpu_soak.py tests/test_gpu_host_files.py -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
tail -5 gpurun_out/${T}_tests.log
timeout 900 python bench.py --stages construct,construct_raw --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/${T}_bench.err
python - <<PY
import json
f="gpurun_out/${T}_bench.json"
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    st=d["stages"]["construct"]
    print("value", d["value"], "ms", st["ms_per_step"], "single", st["single_stream"], st["kernel_ms_per_step"])
    print("e2e", st["e2e"]["value"], st["e2e"]["ms_per_step"], "packed", st["e2e"]["packed_input"])
    r=d["stages"]["construct_raw"]
    print("raw", r["value"], r["ms_per_step"], "e2e", r["e2e"])
except Exception as e:
    print(f, "ERR", e)
PY
python profiles/run/construct_once.py 1000000 2 > gpurun_out/${T}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"insert_words" -s 1 -c 1 -o gpurun_out/${T}_insert python profiles/run/construct_once.py 1000000 2 > gpurun_out/${T}_ncu.log 2>&1
echo ncu rc=$?
