python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_cfg4_v4.json 2> gpurun_out/r2_bench_cfg4_v4.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-fast-path > gpurun_out/r2_bench_cfg5_v4.json 2> gpurun_out/r2_bench_cfg5_v4.err
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2_gputest_v4.log
