"""Developer tool: relative RMS / peak phase error of the GPU against the CPU oracle on small
scenes of every configuration family (one line per scene)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from testkit import synth
from isce3_b200.focus import backproject, last_stats
from oracle import tdbp

oracle = tdbp.best()
tag = sys.argv[1] if len(sys.argv) > 1 else ""
for name, kw in (("c2", dict(pulses=6144, bins=1536, out_lines=24, out_samples=256, n_targets=1)),
                 ("c4", dict(pulses=6144, bins=2048, out_lines=16, out_samples=256, n_targets=1)),
                 ("c5", dict(pulses=6144, bins=1536, out_lines=12, out_samples=200, n_targets=1, taps=8)),
                 ("c5", dict(pulses=6144, bins=1536, out_lines=12, out_samples=200, n_targets=1, taps=16)),
                 # noise-dominated scenes: per-pulse errors do not average out as they do on a point target
                 ("c2", dict(pulses=6144, bins=1536, out_lines=24, out_samples=256, n_targets=1, noise_db=20.0)),
                 ("c4", dict(pulses=6144, bins=2048, out_lines=16, out_samples=256, n_targets=1, noise_db=20.0)),
                 ("c5", dict(pulses=6144, bins=1536, out_lines=12, out_samples=200, n_targets=1, taps=8, noise_db=20.0)),
                 ("c5", dict(pulses=6144, bins=1536, out_lines=12, out_samples=200, n_targets=1, taps=16, noise_db=20.0))):
    sc = synth.make_scene(name, **kw)
    shape = (sc.out_geometry.grid_length, sc.out_geometry.grid_width)
    out = np.zeros(shape, np.complex64)
    ref = np.zeros(shape, np.complex64)
    backproject(out, *sc.backproject_args())
    st = last_stats()
    oracle.backproject(ref, *sc.backproject_args())
    m = np.isfinite(ref)
    rel = np.linalg.norm((out - ref)[m]) / np.linalg.norm(ref[m])
    ph = np.angle(np.sum(out[m] * np.conj(ref[m])))
    print(f"{tag} {name} noise={kw.get('noise_db', -40)} taps={st['taps']} fast={st['used_fast_kernel']} rel_rms {rel:.3e} mean phase {ph:.2e} rad", flush=True)
