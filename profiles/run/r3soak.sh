mkdir -p gpurun_out
T=${1:-r3soak}
for s in stress_search_exit.py:100:9101 stress_first_touch.py:90:9102 stress_other.py:50:9103; do
  IFS=: read name budget seed <<< "$s"
  timeout 300 python tests/soak/$name $budget $seed > gpurun_out/${T}_${name%.py}.log 2>&1; echo "$name rc=$? $(tail -1 gpurun_out/${T}_${name%.py}.log | cut -c1-300)"
done
