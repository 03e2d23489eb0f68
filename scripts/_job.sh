python bench.py --workload cfg2 --steps 10 --warmup 3 --quick-cpu > gpurun_out/r2_bench_cfg2_n1.json 2> gpurun_out/r2_bench_cfg2_n1.err
python bench.py --workload cfg3 --steps 10 --warmup 3 --quick-cpu --n-cpu 100 > gpurun_out/r2_bench_cfg3_n1.json 2> gpurun_out/r2_bench_cfg3_n1.err
python bench.py --workload cfg4 --steps 5 --warmup 3 --quick-cpu --n-cpu 100 > gpurun_out/r2_bench_cfg4_n1.json 2> gpurun_out/r2_bench_cfg4_n1.err
