mkdir -p gpurun_out
T=${1:-r2w}
N=${2:-8}
KWG_GATHER_TRACE=1 NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,P2P,SHM,NET timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --stages search > gpurun_out/${T}_search_${N}gpu.json 2> gpurun_out/${T}_search_${N}gpu.err; echo "bench rc=$?"
grep -c "kwg gather" gpurun_out/${T}_search_${N}gpu.err
grep "kwg gather] root" gpurun_out/${T}_search_${N}gpu.err | tail -6
grep "kwg gather] rank 3" gpurun_out/${T}_search_${N}gpu.err | tail -4
grep -E "via P2P|via SHM|via NET|P2P/|SHM/" gpurun_out/${T}_search_${N}gpu.err | sed 's/.*NCCL INFO//' | sort | uniq -c | sort -rn | head -8
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_search_${N}gpu.json").read().strip().splitlines()[-1])
v=d["stages"]["search"]; print("value %.4g e2e %.4g" % (v["value"], v["e2e"]["value"]))
PY
