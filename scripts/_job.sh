export PROBE_SETTINGS='PLSB_GEMM_STORE_CS=0;PLSB_GEMM_STORE_CS=1'
python scripts/gemm_probe.py 39000 100000 208 200 > gpurun_out/r2_gemm_probe6.txt 2>&1
python scripts/sweep_env.py "" "PLSB_GEMM_STORE_CS=1" >> gpurun_out/r2_gemm_probe6.txt 2>&1
