mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bloom.py tests/test_gpu_soak.py tests/test_gpu_host_files.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
timeout 600 python bench.py --stages construct --no-cpu-baseline > gpurun_out/r2a_construct.json 2> gpurun_out/r2a_construct.err; echo "bench rc=$?"
KWG_COUNT_TWO_LEVEL=1 timeout 600 python bench.py --stages construct --no-cpu-baseline > gpurun_out/r2a_construct_old.json 2> gpurun_out/r2a_construct_old.err; echo "bench old rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2a_construct.json","gpurun_out/r2a_construct_old.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        st=d["stages"]["construct"]
        print(f, d["value"], st["kernel_ms_per_step"], st["e2e"]["value"], st["result"])
    except Exception as e:
        print(f, "ERR", e)
PY
