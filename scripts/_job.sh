timeout 300 python scripts/i8_prof.py 25000 100000 6 0 > gpurun_out/r2_i8_prof15.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_i8_prof15.txt
