set -x
BENCH="python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline"
$BENCH > gpurun_out/r2f_plain.json 2> gpurun_out/r2f_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 700 --csv --log-file gpurun_out/r2f_launches.csv $BENCH > gpurun_out/r2f_ncu_launch.log 2>&1
python tools/prof_gemm.py 1 > gpurun_out/r2f_gemm_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -c 8 -o gpurun_out/r2f_prof_gemm -f python tools/prof_gemm.py 1 > gpurun_out/r2f_ncu_gemm.log 2>&1
ncu -i gpurun_out/r2f_prof_gemm.ncu-rep --page raw --csv > gpurun_out/r2f_prof_gemm_raw.csv 2>/dev/null
ls -la gpurun_out/r2f_*
