mkdir -p gpurun_out
T=${1:-r2r}
ncu --set full --clock-control none --import-source on -k regex:"query_kmers_kernel|query_table_init" -s 6 -c 2 -o gpurun_out/${T}_qk python bench.py --stages search --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_qk.log 2>&1; echo ncu qk rc=$?
tail -3 gpurun_out/${T}_ncu_qk.log
