mkdir -p gpurun_out
T=${1:-r2z}
timeout 1200 python -m pytest tests/test_gpu_bloom.py tests/test_gpu_soak.py tests/test_gpu_fullsize.py tests/test_gpu_packed.py tests/test_gpu_host_files.py -q -m gpu -x > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
tail -4 gpurun_out/${T}_tests.log
timeout 600 python bench.py --stages construct --no-cpu-baseline --steps 6 2> gpurun_out/${T}_bench.err | tee gpurun_out/${T}_bench.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
st = d['stages']['construct']
print('ms', st['ms_per_step'], st['kernel_ms_per_step'], 'e2e', st['e2e']['ms_per_step'], st['e2e']['ascii_input']['ms_per_step'])
"
