mkdir -p gpurun_out
T=r2j
timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_search_gather.py tests/test_gpu_host_files.py -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
tail -4 gpurun_out/${T}_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --stages search --no-cpu-baseline > gpurun_out/${T}_search2.json 2> gpurun_out/${T}_search2.err; echo "bench2 rc=$?"
timeout 600 python bench.py --stages search --no-cpu-baseline > gpurun_out/${T}_search1.json 2> gpurun_out/${T}_search1.err; echo "bench1 rc=$?"
python - <<PY
import json
for f in ("gpurun_out/r2j_search1.json","gpurun_out/r2j_search2.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); st=d["stages"]["search"]
        print(f, st["value"], st["ms_per_step"], st["e2e"]["value"], st["e2e"]["hits"], st["e2e"]["hits_expected_at_least"])
    except Exception as e:
        print(f,"ERR",e)
PY
tail -3 gpurun_out/${T}_search2.err
