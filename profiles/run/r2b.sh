mkdir -p gpurun_out
python profiles/run/construct_once.py 1000000 2 > gpurun_out/r2b_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ft_ -s 2 -c 2 -o gpurun_out/r2b_ft python profiles/run/construct_once.py 1000000 2 > gpurun_out/r2b_ncu.log 2>&1
echo rc=$?; tail -3 gpurun_out/r2b_plain.log; tail -3 gpurun_out/r2b_ncu.log
