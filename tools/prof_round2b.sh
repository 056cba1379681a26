BENCH="python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline --no-side-stream"
$BENCH > gpurun_out/r2_pb_plain.json 2> gpurun_out/r2_pb_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:ln_bwd_kernel -s 2 -c 1 -o gpurun_out/r2_prof_lnbwd2 -f $BENCH > gpurun_out/r2_pb1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ln_fwd_kernel -s 3 -c 1 -o gpurun_out/r2_prof_lnfwd2 -f $BENCH > gpurun_out/r2_pb2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cast_drop_colsum -s 0 -c 1 -o gpurun_out/r2_prof_castdrop -f $BENCH > gpurun_out/r2_pb3.log 2>&1
