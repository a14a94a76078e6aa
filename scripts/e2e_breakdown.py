"""Developer tool: timing breakdown of the one-shot (host-buffer) call on the C2 frame."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from testkit import synth
from isce3_b200.focus import backproject, last_stats

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
sc = synth.make_scene("c2", pulses=16384, bins=int(12288 * scale), out_lines=int(8192 * scale),
                      out_samples=int(8192 * scale), noise_db=False, n_targets=1)
shape = (sc.out_geometry.grid_length, sc.out_geometry.grid_width)
pin = torch.empty(sc.rc.shape, dtype=torch.complex64, pin_memory=True).numpy()
pin[...] = sc.rc
out = torch.empty(shape, dtype=torch.complex64, pin_memory=True).numpy()
args = list(sc.backproject_args())
for host, name in ((pin, "pinned"), (sc.rc, "pageable")):
    args[1] = host
    for it in range(5):
        t = time.perf_counter()
        backproject(out, *args)
        dt = (time.perf_counter() - t) * 1e3
        st = last_stats()
    print(f"{name}: wall {dt:.1f} ms  total {st['ms_total']:.1f}  solve {st['ms_target_solve']:.1f}  h2d {st['ms_h2d']:.1f} "
          f"accumulate {st['ms_accumulate']:.1f}  d2h {st['ms_d2h']:.1f}  launches {st['accumulate_launches']}/{st['total_launches']} "
          f"-> {st['pixel_pulses'] / dt * 1e3:.4g} pp/s", flush=True)
