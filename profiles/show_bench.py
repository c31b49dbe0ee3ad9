"""Print the interesting parts of a bench.py JSON line (profiling aid)."""
import json, sys
d = json.load(open(sys.argv[1]))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], "full", (d.get("full_schedule") or {}).get("value"))
st = d.get("single_trajectory")
if st:
    print("single:", st["scans_per_s"], st["latency_ms"], "launches/frame", st["kernel_launches_per_frame"])
    print("  stages:", {k: round(v, 3) for k, v in st["stage_ms_gpu"].items()})
r = d["roofline"]
print("roofline:", {k: r[k] for k in ("achieved", "frac", "avg_launch_us", "algorithmic_bytes_per_launch", "traffic", "launches")})
print("  batched stage ms:", {k: round(v, 3) for k, v in (r.get("stage_ms_per_frame_batched") or {}).items()})
for k, v in (d.get("knn_throughput") or {}).items():
    print("knn", k, {a: v[a] for a in v if a in ("kernel_ms", "achieved_gbs", "frac_algorithmic", "frac_dram", "traffic", "queries_with_5_neighbours", "problems")} if isinstance(v, dict) else v)
for k, v in (d.get("other_configs") or {}).items():
    print(k, {a: b for a, b in v.items() if a != "what"} if isinstance(v, dict) else v)
print("cpu:", d.get("cpu_baseline"))
print(d["config"]["map_points_in_neighbourhood_last_frame"], d["config"]["fixpoint_skip"]["outer_iterations_run_mean"], d["host"], d["clocks"])
print("knn full-search fraction per outer:", [round(x, 3) for x in (d["config"].get("knn_reuse") or {}).get("full_search_fraction_per_outer_iteration", [])])
