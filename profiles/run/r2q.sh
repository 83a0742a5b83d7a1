mkdir -p gpurun_out
T=${1:-r2q}
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
tail -5 gpurun_out/${T}_tests.log
timeout 1200 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/${T}_bench.err
python - <<PY
import json
f="gpurun_out/${T}_bench.json"
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    for k,v in d["stages"].items():
        print(k, "value %.4g" % v["value"], "ms %.3f" % v["ms_per_step"], "e2e %.4g" % v["e2e"]["value"], "frac", (v.get("roofline") or {}).get("frac"))
    st=d["stages"]["construct"]
    print(st["kernel_ms_per_step"], st["e2e"]["packed_input"]["value"])
    print(d["stages"]["search"]["kernel_ms_per_step"], d["stages"]["construct_raw"]["e2e"])
    print("ingest", d["stages"]["ingest"]["points"]["fastq_gz"]["seconds"], d["stages"]["ingest"]["points"]["plain_reads"]["seconds"])
    print("db_load", {k: d["stages"]["db_load"][k] for k in ("GBps","pinned_h2d_copy_GBps","file_read_alone_GBps")})
    print("cpu", d["cpu_baseline"])
except Exception as e:
    print(f, "ERR", e)
PY
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/${T}_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --stages construct,construct_c5,construct_raw,crc32,transpose,search > gpurun_out/${T}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"query_kmers_kernel|search_count_kernel" -s 6 -c 2 -o gpurun_out/${T}_search python bench.py --stages search --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_qk.log 2>&1; echo ncu qk rc=$?
