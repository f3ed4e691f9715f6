import json,sys
for t in sys.argv[1:]:
    d=json.load(open("gpurun_out/%s.json"%t))
    print(t, round(d["value"]), round(d["roofline"]["frac"],3), round(d["roofline"]["kernel_ms_per_launch"],3), d["config"]["expansion_search"], d["config"]["search_width"], d["config"]["max_iterations"], d["config"]["recall_at_10"], round(d["roofline"]["distance_evals_per_query"]), d["build_s"]["graph"], [(o["search_width"], round(o["ms_per_batch"],2)) for o in d["config"]["operating_points"]])
