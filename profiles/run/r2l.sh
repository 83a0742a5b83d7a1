mkdir -p gpurun_out
T=${1:-r2l}
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
tail -5 gpurun_out/${T}_tests.log
timeout 1200 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/${T}_bench.err
python - <<PY
import json
f="gpurun_out/${T}_bench.json"
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    for k,v in d["stages"].items():
        print(k, "value %.4g" % v["value"], "ms %.3f" % v["ms_per_step"], "e2e %.4g" % v["e2e"]["value"], "frac", (v.get("roofline") or {}).get("frac"))
    st=d["stages"]["construct"]
    print(st["kernel_ms_per_step"], st["e2e"])
    print("ingest", d["stages"]["ingest"]["points"])
    print("db_load", {k: d["stages"]["db_load"][k] for k in ("GBps","pinned_h2d_copy_GBps")})
    print("sweep cross", d["stages"]["sweep"]["l2_to_hbm_crossover"], d["stages"]["sweep"]["seconds"])
    print("cpu", d["cpu_baseline"])
except Exception as e:
    print(f, "ERR", e)
PY
ncu --set full --clock-control none --import-source on -k regex:"transpose_kernel" -s 3 -c 1 -o gpurun_out/${T}_transpose python bench.py --stages transpose --no-cpu-baseline > gpurun_out/${T}_ncu_tr.log 2>&1; echo ncu tr rc=$?
ncu --set full --clock-control none --import-source on -k regex:"search_count_kernel|query_kmers_kernel" -s 6 -c 2 -o gpurun_out/${T}_search python bench.py --stages search --no-cpu-baseline > gpurun_out/${T}_ncu_se.log 2>&1; echo ncu se rc=$?
ncu --set full --clock-control none --import-source on -k regex:"ft_(count|hash|append|resolve|finish)|insert_words" -s 6 -c 6 -o gpurun_out/${T}_construct python profiles/run/construct_once.py 1000000 2 > gpurun_out/${T}_ncu_co.log 2>&1; echo ncu co rc=$?
