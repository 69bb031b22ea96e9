"""Developer tool: one-line summary of a bench.py JSON line read from stdin."""
import json
import sys

d = json.loads(sys.stdin.read().strip().splitlines()[-1])
t = d.get("transform_gemm", {})
r = d["roofline"]
print(f"{d['value'] / 1e9:.2f} Gedges/s  {d['ms_per_step']:.3f} ms/step  gather {sum(v['ms'] for v in r['per_launch'].values()):.3f} ms "
      f"(frac {r['frac']:.2f})  gemm {t.get('ms_per_step', 0):.3f} ms ({t.get('tflops_fp32_equiv', 0):.0f} TF/s)")
print("   ", {k: round(v["ms"], 4) for k, v in r["per_launch"].items()}, {k: round(v["ms"], 4) for k, v in t.get("per_launch", {}).items()})
