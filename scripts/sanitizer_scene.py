"""Small scenes for compute-sanitizer (memcheck / racecheck): one flat-DEM frame through the
fast kernel (steady, per-pulse and aperture-edge runs, row-wavefront one-shot path), one
raster-DEM / Doppler-LUT frame, one generic-kernel frame, one resident plan."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from testkit import synth
from isce3_b200.focus import BackprojectPlan, backproject, last_stats

for name, kw, extra in (
        ("c2", dict(pulses=1536, bins=768, out_lines=12, out_samples=260, n_targets=1), {}),
        ("c5", dict(pulses=1024, bins=768, out_lines=8, out_samples=140, n_targets=1, taps=16), {}),
        ("c4", dict(pulses=640, bins=768, out_lines=8, out_samples=140, n_targets=1, doppler_lut=True), {}),
        ("c1", dict(pulses=512, bins=512, out_lines=8, out_samples=130), dict(force_generic=True))):
    sc = synth.make_scene(name, **kw)
    shape = (sc.out_geometry.grid_length, sc.out_geometry.grid_width)
    out = np.zeros(shape, np.complex64)
    h = np.zeros(shape, np.float32)
    err = backproject(out, *sc.backproject_args(), height=h, batch=200, **extra)
    st = last_stats()
    print(name, "err", err, "fast", st["used_fast_kernel"], "launches", st["total_launches"],
          "finite", bool(np.isfinite(out).all()), flush=True)
sc = synth.make_scene("c2", pulses=1024, bins=768, out_lines=8, out_samples=130, n_targets=1)
with BackprojectPlan(*sc.backproject_args()) as plan:
    plan.execute()
    img = plan.download()
print("plan finite", bool(np.isfinite(img).all()))
