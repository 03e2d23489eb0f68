timeout 300 python scripts/i8_prof.py 25000 100000 6 0 > gpurun_out/r2_i8_prof17.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_i8_prof17.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_gputest_i8_v6.log 2>&1; echo "rc=$?" >> gpurun_out/r2_gputest_i8_v6.log
