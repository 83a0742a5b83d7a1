mkdir -p gpurun_out
timeout 70 python bench.py --stages ingest --no-cpu-baseline 2> gpurun_out/r3i.err > gpurun_out/r3i_ingest.json; echo rc=$?; tail -2 gpurun_out/r3i.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r3i_ingest.json").read().strip().splitlines()[-1]); print(d["stages"]["ingest"]["points"])
PY
