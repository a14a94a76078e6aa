import sys, os
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from isce3_b200 import synth
from isce3_b200.container import RadarGeometry
from isce3_b200.core import LUT2d
from isce3_b200.focus import backproject, last_stats
from oracle import tdbp
sc = synth.make_scene("c2", pulses=3000, bins=1024, out_lines=5000, out_samples=4, n_targets=1, out_prf_ratio=1.0)
g = sc.out_geometry.radar_grid
g.prf = g.prf / 250.0
g.length = 20
sc.out_geometry = RadarGeometry(g, sc.out_geometry.orbit, LUT2d())
shape = (20, 4)
ref = np.zeros(shape, np.complex64)
tdbp.best().backproject(ref, *sc.backproject_args())
for trial in range(3):
    for batch in (97, 100000, 1024):
        for gen in (False, True):
            out = np.zeros(shape, np.complex64)
            backproject(out, *sc.backproject_args(), batch=batch, force_generic=gen)
            st = last_stats()
            nanrows = np.where(np.isnan(out.real).any(axis=1))[0]
            m = np.isfinite(out.real)
            rel = np.linalg.norm((out - ref)[m]) / np.linalg.norm(ref[m])
            print(trial, batch, "generic" if gen else "fast", "nan rows", nanrows.tolist(), "rel", f"{rel:.2e}", "pulses", st["pulse_first"], st["pulse_last"], flush=True)
