"""Developer tool: one-line summary of a bench.py JSON line read from stdin."""
import json
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else ""
line = [l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]
d = json.loads(line)
r = d.get("roofline") or {}
print(f"{tag}: value {d['value']:.4g} pp/s  kernel {r.get('pp_per_s_kernel', 0):.4g} pp/s  "
      f"fp32 frac {r.get('frac', 0):.3f}  kernel_ms {r.get('kernel_ms_per_step', 0):.1f}  "
      f"step_ms {d['ms_per_step']:.1f}  e2e {d['e2e']['value']:.4g}  solve_ms {d.get('target_solve_ms_per_step', 0):.1f} "
      f"clk {d['clocks']['sm_mhz'] if d.get('clocks') else None} {d['clocks']['reasons'] if d.get('clocks') else None}")
