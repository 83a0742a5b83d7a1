mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_host_files.py tests/test_gpu_pipeline_vs_reference.py -q -m gpu -x > gpurun_out/r3h_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r3h_tests.log
timeout 100 python bench.py --stages ingest --no-cpu-baseline 2>/dev/null > gpurun_out/r3h_ingest.json; python - <<PY
import json
d=json.loads(open("gpurun_out/r3h_ingest.json").read().strip().splitlines()[-1]); print(d["stages"]["ingest"]["points"])
PY
