timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest_i8_v8.log 2>&1; echo "rc=$?" >> gpurun_out/r2_gputest_i8_v8.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_i8.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_smoke_i8.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_cfg5_i8_v4.json 2> gpurun_out/r2_bench_cfg5_i8_v4.err; echo "rc=$?" >> gpurun_out/r2_bench_cfg5_i8_v4.err
