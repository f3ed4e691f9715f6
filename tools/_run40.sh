timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:graph_search_cta_kernel --launch-skip 25 --launch-count 1 -o gpurun_out/r40_k4b_batch1 -f python tools/_batch1_probe.py > gpurun_out/r40_ncu.log 2>&1
exit 0
