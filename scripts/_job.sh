python -m pytest tests/test_gpu_large_k.py tests/test_gpu_splithalf.py tests/test_gpu_parity.py -x -q -m gpu --durations=8 2>&1 | tail -40 > gpurun_out/r2_largek_v2.log
