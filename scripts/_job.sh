timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest_i8_v1.log 2>&1; echo "rc=$?" >> gpurun_out/r2_gputest_i8_v1.log
