python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_gputest_v2.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_cfg5_v2.json 2> gpurun_out/r2_bench_cfg5_v2.err
python bench.py --workload cfg2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_cfg2_v2.json 2> gpurun_out/r2_bench_cfg2_v2.err
