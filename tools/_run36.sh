timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --rows 10000000 --storage bf16 --clusters 2560 --no-cpu-baseline --steps 10 > gpurun_out/r36_c3_n8.log 2>&1
exit 0
