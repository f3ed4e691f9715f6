timeout 900 python -m pytest tests/test_gpu_tc.py -x -q -m gpu -s 2>&1 | tail -40 > gpurun_out/r34_tc_tests.log
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -12 > gpurun_out/r34_parity.log
timeout 300 python tools/k4_sweep.py --n 1000000 --dim 768 --batch 10000 --storage f32 --exact-only > gpurun_out/r34_exact_f32.log 2>&1
timeout 300 python tools/k4_sweep.py --n 1000000 --dim 768 --batch 10000 --storage bf16 --exact-only > gpurun_out/r34_exact_bf16.log 2>&1
VSB_CERT_KP16=32 timeout 300 python tools/k4_sweep.py --n 1000000 --dim 768 --batch 10000 --storage bf16 --exact-only > gpurun_out/r34_exact_bf16_kp32.log 2>&1
exit 0
