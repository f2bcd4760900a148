import json, sys
d = [json.loads(x) for x in open(sys.argv[1]) if x.strip().startswith("{")][-1]
r = d["roofline"]
print("value", round(d["value"]), "e2e", round((d.get("e2e") or {}).get("value", 0)), "launches", d.get("gpu_launches"), "ms/step", round(d["ms_per_step"], 3))
print("dominant", r.get("dominant_kernel"), "frac", round(r["frac"], 4), "avg_launch_us", round(r["avg_launch_us"], 1), "fpl", r.get("frames_per_launch"))
for k, v in r["all_timed_kernels"].items():
    print(" ", k, "launch_us", round(v["avg_launch_us"], 1), "fpl", v.get("frames_per_launch"), "us/frame", round(v.get("us_per_frame", 0), 2), "frac", round(v["frac"], 4), "span_us", round(v["span_in_timed_region_us"], 1))
print("untimed pass per-frame us:", {k: round(v / d["config"]["frames_per_gpu_per_step"] * 1000, 1) for k, v in r["kernel_ms_per_step_untimed_pass"].items() if v})
c = d["config"]
for k in ("worst_case_distribution", "small_frame", "strong_scaling", "config5"):
    if c.get(k) and "value" in c[k]:
        print(k, round(c[k]["value"]), {kk: (round(vv, 2) if isinstance(vv, float) else vv) for kk, vv in c[k].items() if kk in ("us_per_frame", "ratio_to_natural_frame_time", "entropy_form", "ms")})
print("pipeline", d.get("pipeline") or r.get("pipeline"))
print("clocks", d.get("clocks"))
