mkdir -p gpurun_out
T=${1:-r2o}
timeout 1500 python -m pytest tests/test_gpu_bloom.py tests/test_gpu_search.py tests/test_gpu_search_gather.py tests/test_gpu_host_files.py tests/test_gpu_packed.py tests/test_reference_shim.py -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
tail -5 gpurun_out/${T}_tests.log
timeout 400 python tests/soak/stress_first_touch.py 150 777 > gpurun_out/${T}_soak.log 2>&1; echo "soak rc=$?"; tail -2 gpurun_out/${T}_soak.log
timeout 900 python bench.py --stages search,sweep --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/${T}_bench.err
python - <<PY
import json
f="gpurun_out/${T}_bench.json"
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    st=d["stages"]
    print("search", st["search"]["value"], st["search"]["ms_per_step"], st["search"]["e2e"]["value"], st["search"]["kernel_ms_per_step"], st["search"]["roofline"]["frac"])
    for r in st["sweep"]["raw_construction"]:
        if r["num_hash"]==3 and r["log2_len"] in (26,29): print(r["k"], r["log2_len"], "%.3g"%r["kmer_inserts_per_s"], r["parity"])
except Exception as e:
    print(f, "ERR", e)
PY
