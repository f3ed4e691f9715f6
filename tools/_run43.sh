timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r43_tests.log
VSB200_LIB=$PWD/vector-store_b200/libvsb200_prof.so timeout 500 python tools/_batch1_probe.py > gpurun_out/r43_k4b_phases.log 2>&1
timeout 500 python tools/latency_breakdown.py > gpurun_out/r43_latency.log 2>&1
exit 0
