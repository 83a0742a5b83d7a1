mkdir -p gpurun_out
for w in 4 5 3; do
  timeout 100 python bench.py --stages construct --no-cpu-baseline --steps 9 --e2e-workers $w 2>/dev/null > gpurun_out/r3w_$w.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/r3w_$w.json").read().strip().splitlines()[-1]); e=d["stages"]["construct"]["e2e"]
print("workers", e["workers_per_gpu"], "packed ms", round(e["ms_per_step"],3), "ascii ms", round(e["ascii_input"]["ms_per_step"],3), "device ms", round(d["stages"]["construct"]["ms_per_step"],3))
PY
done
