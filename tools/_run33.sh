set -x
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q -m gpu -s 2>&1 | tail -30 > gpurun_out/r33_tc_tests.log
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r33_parity.log
for kp in 64 128 256; do
VSB_CERT_KP=$kp timeout 300 python tools/k4_sweep.py --n 1000000 --dim 768 --batch 10000 --storage f32 --exact-only > gpurun_out/r33_exact_f32_kp$kp.log 2>&1
done
VSB_DISABLE_CERT=1 timeout 300 python tools/k4_sweep.py --n 1000000 --dim 768 --batch 10000 --storage f32 --exact-only > gpurun_out/r33_exact_f32_simt.log 2>&1
tail -3 gpurun_out/r33_*.log
