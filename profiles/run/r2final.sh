mkdir -p gpurun_out
T=${1:-r2final}
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${T}_smoke.log
bash profiles/run/r2q.sh $T
