"""Print the essentials of bench.py JSON lines: python tools/show_bench.py file.json [...]"""
import json
import sys

for f in sys.argv[1:]:
    ls = [l for l in open(f) if l.startswith("{")]
    if not ls:
        print(f, "NO JSON LINE:", open(f).read()[-300:])
        continue
    j = json.loads(ls[-1])
    e = j.get("e2e") or {}
    print(f"{f}: N={j.get('n_gpus')} ms/step {j['ms_per_step']:.3f} (min {j.get('ms_per_step_min', 0):.3f}) {j['value'] / 1e9:.2f} G  e2e {e.get('ms_per_step', 0):.3f} ms (min {e.get('ms_per_step_min', 0):.2f} med {e.get('ms_per_step_median', 0):.2f} max {e.get('ms_per_step_max', 0):.2f})  "
          f"launches/step {j['gpu_launches'] / j['steps']:.1f}  parity {j.get('parity_check')}  checksum {j.get('state_checksum', {}).get('value', [None] * 2)[:2]}")
    r = j["roofline"]
    print("    phases", {k: round(v["ms_per_step"], 3) for k, v in r["phases"].items()}, "kernel frac", round(r["frac"], 3), "step frac", round(r["step"]["frac"], 3),
          "traffic/alg", round(r["traffic"] / (r["alg_bytes_per_particle"] * j["config"]["particles_mean"] / j.get("n_gpus", 1)), 2) if r.get("traffic") else None)
    for k, v in j.get("extra", {}).items():
        if "error" in v:
            print("    extra", k, "ERROR", v["error"])
        else:
            print(f"    extra {k}: ms/step {v['ms_per_step']:.3f} {v['value'] / 1e9:.2f} G particles {v['config']['particles_mean']:.0f} step frac {v['roofline']['step']['frac']:.3f}")
    if "cpu_baseline" in j:
        print("    cpu_baseline", j["cpu_baseline"].get("value"), "cores", j["cpu_baseline"].get("cores"))
