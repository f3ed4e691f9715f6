timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r39_tests.log
timeout 500 python tools/latency_breakdown.py > gpurun_out/r39_latency.log 2>&1
exit 0
