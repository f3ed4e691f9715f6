timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r41_tests.log
timeout 500 python tools/latency_breakdown.py > gpurun_out/r41_latency.log 2>&1
exit 0
