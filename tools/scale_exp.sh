run() { tag=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 --no-extras "$@" > gpurun_out/r2_scale_$tag.json 2> gpurun_out/r2_scale_$tag.err; }
run ctas8 --nccl-ctas 8
run ctas16 --nccl-ctas 16
MMOE_DYNAMIC_TILES=1 run dyn8 --nccl-ctas 8
