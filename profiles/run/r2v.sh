mkdir -p gpurun_out
T=${1:-r2v}
N=${2:-8}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N > gpurun_out/${T}_bench_${N}gpu.json 2> gpurun_out/${T}_bench_${N}gpu.err; echo "bench rc=$?"
tail -3 gpurun_out/${T}_bench_${N}gpu.err
python - <<PY
import json
f="gpurun_out/${T}_bench_${N}gpu.json"
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    for k,v in d["stages"].items():
        print(k, "value %.4g" % v["value"], "ms %.3f" % v["ms_per_step"], "e2e %.4g" % v["e2e"]["value"])
    print(d["stages"]["construct"]["e2e"]); print(d["stages"]["construct_raw"]["e2e"]); print(d["stages"]["search"]["e2e"]); print(d["clocks"])
except Exception as e:
    print(f, "ERR", e)
PY
