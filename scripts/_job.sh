timeout 200 python scripts/i8_gemm_check.py quick > gpurun_out/r2_i8_check8.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_i8_check8.txt
timeout 300 python scripts/i8_prof.py 25000 100000 6 0,8 > gpurun_out/r2_i8_prof7.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_i8_prof7.txt
PLSB_I8_NP0=3 timeout 300 python scripts/i8_prof.py 25000 100000 6 0 >> gpurun_out/r2_i8_prof7.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_i8_prof7.txt
PLSB_I8_NP0=3 timeout 200 python scripts/i8_gemm_check.py quick >> gpurun_out/r2_i8_check8.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_i8_check8.txt
