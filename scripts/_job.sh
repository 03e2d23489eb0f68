python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py -x -q -m gpu --durations=6 -k "scale or largest or 200_plus or full_n" 2>&1 | tail -14 > gpurun_out/r2_scale_v3.log
