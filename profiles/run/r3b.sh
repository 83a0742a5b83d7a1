mkdir -p gpurun_out
T=${1:-r3b}
timeout 300 python -m pytest tests/test_gpu_bloom.py tests/test_gpu_packed.py -q -m gpu -x -k "counting_mode or dup_kmers or staged or packed or filters_beyond or golden or finalize" > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
tail -3 gpurun_out/${T}_tests.log
timeout 120 python tests/soak/stress_first_touch.py ${SOAK:-30} 777 > gpurun_out/${T}_soak.log 2>&1; echo "soak rc=$? $(tail -1 gpurun_out/${T}_soak.log | cut -c1-120)"
timeout 600 python bench.py --stages construct --no-cpu-baseline --steps 6 2> gpurun_out/${T}_bench.err | tee gpurun_out/${T}_bench.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
st = d['stages']['construct']
print('ms', st['ms_per_step'], st['kernel_ms_per_step'], 'e2e', st['e2e']['ms_per_step'], st['e2e']['ascii_input']['ms_per_step'])
"
tail -3 gpurun_out/${T}_bench.err
