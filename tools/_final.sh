timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r49_tests.log
timeout 900 python bench.py > gpurun_out/r49_bench_n1.log 2>&1
timeout 600 python bench.py --traversal i8 --no-cpu-baseline > gpurun_out/r49_bench_n1_i8.log 2>&1
exit 0
