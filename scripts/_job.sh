python -m pytest tests/test_gpu_primitives.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4 > gpurun_out/r2_smallk_test3.log
python scripts/sweep_env.py "" "PLSB_GP_CH=192" > gpurun_out/r2_sweep_gps2.txt 2>&1
python scripts/sweep_env.py --workload cfg2 "" "PLSB_SMALL_K=0" >> gpurun_out/r2_sweep_gps2.txt 2>&1
python scripts/sweep_env.py --workload cfg3 "" "PLSB_SMALL_K=0" >> gpurun_out/r2_sweep_gps2.txt 2>&1
