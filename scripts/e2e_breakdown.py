"""Developer tool: timing breakdown of the one-shot (host-buffer) call.
Usage: e2e_breakdown.py [config=c2] [pinned|pageable|both] [iterations]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from testkit import synth
from isce3_b200.focus import BackprojectPlan, backproject, last_stats

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
mode = sys.argv[2] if len(sys.argv) > 2 else "both"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 4
kw = {"n_targets": 9} if cfg == "c4" else {}
sc = synth.make_scene(cfg, **kw)
shape = (sc.out_geometry.grid_length, sc.out_geometry.grid_width)
args = list(sc.backproject_args())
with BackprojectPlan(*args) as plan:
    for _ in range(3):
        plan.execute()
    st = plan.stats()
    print(f"resident: total {st['ms_total']:.1f} solve {st['ms_target_solve']:.1f} accumulate {st['ms_accumulate']:.1f}", flush=True)
hosts = []
if mode in ("pinned", "both"):
    import torch
    pin = torch.empty(sc.rc.shape, dtype=torch.complex64, pin_memory=True).numpy()
    pin[...] = sc.rc
    out = torch.empty(shape, dtype=torch.complex64, pin_memory=True).numpy()
    hosts.append((pin, out, "pinned"))
if mode in ("pageable", "both"):
    hosts.append((sc.rc, np.empty(shape, np.complex64), "pageable"))
for host, out, name in hosts:
    args[1] = host
    for it in range(iters):
        t = time.perf_counter()
        backproject(out, *args)
        dt = (time.perf_counter() - t) * 1e3
        st = last_stats()
    print(f"{name}: wall {dt:.1f} ms  total {st['ms_total']:.1f}  solve {st['ms_target_solve']:.1f}  h2d {st['ms_h2d']:.1f} "
          f"accumulate {st['ms_accumulate']:.1f}  d2h {st['ms_d2h']:.1f}  launches {st['accumulate_launches']}/{st['total_launches']} "
          f"-> {st['pixel_pulses'] / dt * 1e3:.4g} pp/s", flush=True)
