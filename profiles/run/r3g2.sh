mkdir -p gpurun_out
T=${1:-r3g2}
nvidia-smi -L | head -3
timeout 300 python -m pytest tests/test_gpu_search_gather.py tests/test_gpu_host_files.py -q -m gpu -x -k "gather or devices or slab or kwage" > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
tail -4 gpurun_out/${T}_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --stages construct,search --no-cpu-baseline > gpurun_out/${T}_bench_2gpu.json 2> gpurun_out/${T}_bench_2gpu.err; echo "bench rc=$?"
tail -3 gpurun_out/${T}_bench_2gpu.err
python - <<PY
import json
f="gpurun_out/${T}_bench_2gpu.json"
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    for k,v in d["stages"].items():
        print(k, "value %.4g" % v["value"], "ms %.3f" % v["ms_per_step"], "e2e %.4g" % v["e2e"]["value"])
    print(d["stages"]["search"]["e2e"])
except Exception as e:
    print(f, "ERR", e)
PY
