timeout 600 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_gputest_i8_v9.log 2>&1; echo "rc=$?" >> gpurun_out/r2_gputest_i8_v9.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_cfg5_i8_v5.json 2> gpurun_out/r2_bench_cfg5_i8_v5.err; echo "rc=$?" >> gpurun_out/r2_bench_cfg5_i8_v5.err
