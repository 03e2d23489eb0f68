timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest_final.log 2>&1; echo "rc=$?" >> gpurun_out/r2_gputest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_final.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_smoke_final.txt
