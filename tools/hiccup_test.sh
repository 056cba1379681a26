for i in 1 2 3; do
  BENCH_NO_SAMPLER=1 python bench.py --steps 30 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('nosampler', d['ms_per_step'], d['ms_per_step_min_median_max'], d['slowest_step'], d['host_ms_max_step'])"
  python bench.py --steps 30 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('sampler  ', d['ms_per_step'], d['ms_per_step_min_median_max'], d['slowest_step'], d['host_ms_max_step'], d['clocks'])"
done
