timeout 150 python scripts/i8_gemm_check.py quick > gpurun_out/r2_i8_check.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_i8_check.txt
