timeout 600 python tools/c2_sweep.py > gpurun_out/r37_c2_sweep.log 2>&1
exit 0
