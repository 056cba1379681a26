set -x
BENCH="python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline"
$BENCH > gpurun_out/r2_p_plain.json 2> gpurun_out/r2_p_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:ln_bwd_kernel -s 20 -c 1 -o gpurun_out/r2_prof_lnbwd -f $BENCH > gpurun_out/r2_p1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_tc_bwd -s 4 -c 1 -o gpurun_out/r2_prof_attnbwd -f $BENCH > gpurun_out/r2_p2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ln_fwd_kernel -s 20 -c 1 -o gpurun_out/r2_prof_lnfwd -f $BENCH > gpurun_out/r2_p3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_tc_fwd -s 4 -c 1 -o gpurun_out/r2_prof_attnfwd -f $BENCH > gpurun_out/r2_p4.log 2>&1
HEAD="python tools/prof_head.py 65536 bf16 2"
$HEAD > gpurun_out/r2_p_head.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:head_ -s 4 -c 2 -o gpurun_out/r2_prof_head -f $HEAD > gpurun_out/r2_p5.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail
