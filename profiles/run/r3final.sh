mkdir -p gpurun_out
T=${1:-r3final}
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${T}_smoke.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
tail -4 gpurun_out/${T}_tests.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/${T}_bench.err
python - <<PY
import json
f="gpurun_out/${T}_bench.json"
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    for k,v in d["stages"].items():
        print(k, "value %.4g" % v["value"], "ms %.3f" % v["ms_per_step"], "e2e %.4g" % v["e2e"]["value"], "frac", (v.get("roofline") or {}).get("frac"))
    st=d["stages"]["construct"]
    print(st["kernel_ms_per_step"], st["e2e"]["ms_per_step"])
    print("cpu", d["cpu_baseline"])
except Exception as e:
    print(f, "ERR", e)
PY
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/${T}_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --stages construct,construct_c5,construct_raw,crc32,transpose,search > gpurun_out/${T}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"ft_(count|hash|append|resolve|finish)|fold_touched|insert_hash" -s 7 -c 7 -o gpurun_out/${T}_construct python profiles/run/construct_once.py 1000000 2 > gpurun_out/${T}_ncu_co.log 2>&1; echo ncu co rc=$?
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"query_kmers_kernel|search_count_kernel" -s 6 -c 2 -o gpurun_out/${T}_search python bench.py --stages search --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_se.log 2>&1; echo ncu se rc=$?
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"search_count_kernel<3, true>" -s 3 -c 1 -o gpurun_out/${T}_search_exit python bench.py --stages search --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_sx.log 2>&1; echo ncu sx rc=$?
ls -la gpurun_out/${T}_*.ncu-rep
