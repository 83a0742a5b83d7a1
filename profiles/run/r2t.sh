mkdir -p gpurun_out
T=${1:-r2t}
for V in "" "-DKWG_FLUSH_SPREAD=0"; do
  KWG_NVCC_EXTRA="$V" python -c "from kwage_b200 import build; build.build(force=True)" > /dev/null 2>&1
  echo "== variant [$V]"
  timeout 600 python bench.py --stages construct --no-cpu-baseline --steps 6 2> /dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
st = d['stages']['construct']
print('ms', st['ms_per_step'], st['kernel_ms_per_step'], 'e2e', st['e2e']['ms_per_step'])
"
done
python -c "from kwage_b200 import build; build.build(force=True)" > /dev/null 2>&1
timeout 900 python -m pytest tests/test_gpu_bloom.py tests/test_gpu_soak.py -q -m gpu -k "counting or soak or golden or first" 2>&1 | tail -3
