"""Developer tool: time plan.execute() of the scaled C2 frame for one library build."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from testkit import synth
from isce3_b200.focus import BackprojectPlan

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
tag = sys.argv[2] if len(sys.argv) > 2 else ""
kw = dict(pulses=16384, bins=int(12288 * scale), out_lines=int(8192 * scale), out_samples=int(8192 * scale), noise_db=False, n_targets=1)
if len(sys.argv) > 3:
    kw["taps"] = int(sys.argv[3])
name = sys.argv[4] if len(sys.argv) > 4 else "c2"
if name == "c5":
    kw = dict(pulses=16384, bins=4096, out_lines=int(2048 * scale), out_samples=2048, noise_db=False, n_targets=1,
              taps=kw.get("taps", 16))
sc = synth.make_scene(name, **kw)
with BackprojectPlan(*sc.backproject_args()) as plan:
    plan.execute()
    ms = []
    for _ in range(3):
        plan.execute()
        st = plan.stats()
        ms.append(st["ms_accumulate"])
    pp = st["pixel_pulses"]
    best = min(ms)
    taps = st["taps"]
    print(f"{tag}: variant={st["fast_variant"]} kernel {best:.1f} ms  {pp / best * 1e3:.4g} pp/s  frac(73.9T) {(34 + 10 * taps) * pp / best * 1e3 / 73.9e12:.3f} fast={st['used_fast_kernel']} all={['%.1f' % m for m in ms]}", flush=True)
