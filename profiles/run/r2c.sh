mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bloom.py tests/test_gpu_soak.py tests/test_gpu_host_files.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c_tests.log
tail -4 gpurun_out/r2c_tests.log
timeout 600 python bench.py --stages construct --no-cpu-baseline > gpurun_out/r2c_construct.json 2> gpurun_out/r2c_construct.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2c_construct.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        st=d["stages"]["construct"]
        print(f, d["value"], st["kernel_ms_per_step"], st["e2e"]["value"], st["result"])
    except Exception as e:
        print(f, "ERR", e)
PY
python profiles/run/construct_once.py 1000000 2 > gpurun_out/r2c_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ft_ -s 2 -c 2 -o gpurun_out/r2c_ft python profiles/run/construct_once.py 1000000 2 > gpurun_out/r2c_ncu.log 2>&1
echo ncu rc=$?
