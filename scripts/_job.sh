python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2_gputest_v5.log
