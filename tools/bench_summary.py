"""Developer tool: short summary of a bench.py JSON line read from stdin."""
import json
import sys

try:
    d = json.loads(sys.stdin.read().strip().splitlines()[-1])
except Exception as e:  # empty / failed run
    print("no bench line:", e)
    sys.exit(0)
t = d.get("transform_gemm", {})
r = d["roofline"]
print(f"{d['value'] / 1e9:.2f} Gedges/s  {d['ms_per_step']:.3f} ms/step  gather {sum(v['ms'] for v in r['per_launch'].values()):.3f} ms "
      f"(frac {r['frac']:.2f} of {r['bound']} peak {r['peak']:.0f} GB/s; hbm-model {r.get('hbm_model_frac') or 0:.2f}, "
      f"compulsory {r.get('hbm_compulsory_frac') or 0:.2f})  gemm {t.get('ms_per_step', 0):.3f} ms ({t.get('tflops_fp32_equiv', 0):.0f} TF/s, "
      f"frac {t.get('frac', 0):.2f})")
print("   ", {k: round(v["ms"], 4) for k, v in r["per_launch"].items()}, {k: round(v["ms"], 4) for k, v in t.get("per_launch", {}).items()})
print("    probe GB/s:", {k: round(v.get("l2_probe_gbs", 0)) for k, v in r["per_launch"].items()})
if "e2e" in d:
    e = d["e2e"]
    print(f"    e2e {e['value'] / 1e9:.2f} Gedges/s {e['ms_per_step']:.3f} ms  {e.get('breakdown')}")
if "collective" in d:
    print("    collective:", d["collective"], "scaling:", d["scaling"], "n_gpus:", d["n_gpus"])
