mkdir -p gpurun_out
T=${1:-r2soak}
for s in stress_first_touch.py:120:9001 stress_medium.py:100:9002 stress_levels.py:80:9003 stress_other.py:80:9004 stress_thresholds.py:60:9005; do
  IFS=: read name budget seed <<< "$s"
  timeout 400 python tests/soak/$name $budget $seed > gpurun_out/${T}_${name%.py}.log 2>&1; echo "$name rc=$? $(tail -1 gpurun_out/${T}_${name%.py}.log | cut -c1-100)"
done
