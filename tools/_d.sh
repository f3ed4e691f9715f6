timeout 400 python tools/diag_zero_recall.py > gpurun_out/r58_diag.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "reachable or recall or streaming or hybrid or snapshot" -s 2>&1 | tail -15 > gpurun_out/r58_tests.log
exit 0
