mkdir -p gpurun_out
T=${1:-r2p}
timeout 1500 python -m pytest tests/test_gpu_bloom.py tests/test_gpu_search.py tests/test_gpu_search_gather.py tests/test_gpu_host_files.py tests/test_gpu_packed.py tests/test_reference_shim.py tests/test_gpu_pipeline_vs_reference.py -q -m gpu > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
tail -5 gpurun_out/${T}_tests.log
timeout 900 python bench.py --stages search --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/${T}_bench.err
python - <<PY
import json
f="gpurun_out/${T}_bench.json"
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    st=d["stages"]
    print("search", st["search"]["value"], st["search"]["ms_per_step"], st["search"]["e2e"], st["search"]["kernel_ms_per_step"], st["search"]["roofline"]["frac"])
except Exception as e:
    print(f, "ERR", e)
PY
