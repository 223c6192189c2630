"""Print the essentials of a bench.py JSON line:  python profiles/show_bench.py gpurun_out/bench.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"{d['metric']}: {d['value']:.4g} {d['unit']} on {d['n_gpus']} GPU(s), {d['ms_per_step']:.4f} ms/step; roofline frac "
      f"{d['roofline']['frac']:.3f}; launches {d.get('gpu_launches')}; clocks {d.get('clocks')}")
print("e2e:", d.get("e2e"))
print("cpu:", d.get("cpu_baseline"))
print("verify:", d.get("verify"))
for s in d.get("other_configs", []):
    if "error" in s:
        print("  ERROR", s)
        continue
    print(f"  {s['config']:4s} {s['systems_per_s']:.4g} sys/s  hbm_frac {s['hbm_frac']:.3f}  ms/step {s['ms_per_step']:.4f}  verified "
          f"{s.get('verified_systems')} mismatches {s.get('mismatches')}  | {s['workload'][:100]}")
    for k in ("e2e", "per_n", "max_rel_residual_long_double", "round_trip_rel_err_max", "median_launch_ms"):
        if k in s:
            print(f"        {k}: {s[k]}")
