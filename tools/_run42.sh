VSB200_LIB=$PWD/vector-store_b200/libvsb200_prof.so timeout 500 python tools/_batch1_probe.py > gpurun_out/r42_k4b_phases.log 2>&1
timeout 500 python tools/c2_sweep.py --ks 100 --batches 1,256,10000 > gpurun_out/r42_k100.log 2>&1
exit 0
