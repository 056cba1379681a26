run() { tag=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 --no-extras "$@" > gpurun_out/r2_scale_$tag.json 2> gpurun_out/r2_scale_$tag.err; }
MMOE_CROSS_STAGED=0 run ddp_nostage --exchange ddp
MMOE_NATIVE_DEFER=1 run native_defer --exchange native
