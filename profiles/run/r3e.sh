mkdir -p gpurun_out
T=${1:-r3e}
timeout 600 python -m pytest tests/test_gpu_search.py tests/test_gpu_search_gather.py tests/test_gpu_host_files.py tests/test_reference_shim.py tests/test_gpu_bloom.py -q -m gpu -x -k "search or checkpoint or finalize or make_bloom_filter or shim or reference" > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
tail -5 gpurun_out/${T}_tests.log
timeout 600 python bench.py --stages search --no-cpu-baseline --steps 5 2> gpurun_out/${T}_bench.err | tee gpurun_out/${T}_bench.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
st = d['stages']['search']
print('ms', st['ms_per_step'], st.get('kernel_ms_per_step'), 'e2e', json.dumps(st['e2e']))
"
tail -3 gpurun_out/${T}_bench.err
KWG_SEARCH_NO_EXIT=1 timeout 600 python bench.py --stages search --no-cpu-baseline --steps 5 2> gpurun_out/${T}_bench_noexit.err | tee gpurun_out/${T}_bench_noexit.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
st = d['stages']['search']
print('NO_EXIT ms', st['ms_per_step'], 'e2e', json.dumps(st['e2e']))
"
