"""Developer tool: run the reference CUDA backprojection (reduced harness) next to ours."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from testkit import synth
from isce3_b200.focus import backproject, last_stats
from oracle import tdbp

rc = tdbp.ref_cuda()
def run(name, **kw):
    sc = synth.make_scene(name, **kw)
    shape = (sc.out_geometry.grid_length, sc.out_geometry.grid_width)
    ours = np.zeros(shape, np.complex64)
    backproject(ours, *sc.backproject_args())
    pp = last_stats()["pixel_pulses"]
    ref = np.zeros(shape, np.complex64)
    for it in range(2):
        t = time.perf_counter()
        e = rc.backproject(ref, *sc.backproject_args())
        dt = time.perf_counter() - t
    cpu = np.zeros(shape, np.complex64)
    if shape[0] * shape[1] * 4000 < 3e9:
        tdbp.best().backproject(cpu, *sc.backproject_args())
        rel_cpu = np.linalg.norm(ref - cpu) / np.linalg.norm(cpu)
    else:
        rel_cpu = float("nan")
    rel = np.linalg.norm(ref - ours) / np.linalg.norm(ours)
    print(f"{name} {kw}: refcuda err={e} {dt*1e3:.1f} ms -> {pp/dt:.4g} pp/s ; rel(refcuda vs ours) {rel:.2e} rel(refcuda vs cpu) {rel_cpu:.2e}", flush=True)

run("c1", pulses=2048, bins=4096, out_lines=64, out_samples=512)
run("c2", pulses=6144, bins=2048, out_lines=64, out_samples=512, n_targets=1)
run("c2", pulses=16384, bins=12288, out_lines=512, out_samples=8192, n_targets=1, noise_db=False)
