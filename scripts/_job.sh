set -x
python bench.py --workload cfg5 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-fast-path > gpurun_out/r2_cfg5_base.json 2> gpurun_out/r2_cfg5_base.err
python scripts/cublas_shape_probe.py 25000 100000 208 > gpurun_out/r2_cublas_probe.txt 2>&1
python scripts/cublas_shape_probe.py 25000 100000 200 >> gpurun_out/r2_cublas_probe.txt 2>&1
python scripts/cublas_shape_probe.py 8192 8192 8192 >> gpurun_out/r2_cublas_probe.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_cfg5_launches_base.csv python bench.py --workload cfg5 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-fast-path > gpurun_out/ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:xcov_gemm -s 4 -c 4 -o gpurun_out/r2_cfg5_gemm_base python bench.py --workload cfg5 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-fast-path --workspace-gib 8 > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out
