VSB_BUILD_TIMING=1 timeout 1100 python bench.py --n 10000000 --storage bf16 --clusters 2560 --no-cpu-baseline --steps 10 > gpurun_out/r35_c3_n1.log 2>&1
nvidia-smi --query-gpu=memory.used --format=csv >> gpurun_out/r35_c3_n1.log
exit 0
