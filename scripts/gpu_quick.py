"""Quick GPU sanity run (developer tool): parity of generic and fast kernels vs oracle."""
import sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from testkit import synth
from isce3_b200.focus import backproject, last_stats, measure_peaks
from oracle import tdbp

def run(name, **kw):
    sc = synth.make_scene(name, **kw)
    shape = (sc.out_geometry.grid_length, sc.out_geometry.grid_width)
    ref = np.zeros(shape, np.complex64); href = np.zeros(shape, np.float32)
    t = time.time(); o = tdbp.best(); oe = o.backproject(ref, *sc.backproject_args(), height=href); tc = time.time() - t
    for generic in (True, False):
        out = np.zeros(shape, np.complex64); h = np.zeros(shape, np.float32)
        t = time.time()
        e = backproject(out, *sc.backproject_args(), batch=256, height=h, force_generic=generic)
        tg = time.time() - t
        st = last_stats()
        m = np.isfinite(ref)
        rel = np.linalg.norm((out - ref)[m]) / np.linalg.norm(ref[m])
        print(f"{name} {kw} generic={generic}: err {e}/{oe} rel {rel:.3e} nanmatch {np.array_equal(np.isnan(out), np.isnan(ref))} "
              f"hmax {np.nanmax(np.abs(h-href)):.2e} fast={st['used_fast_kernel']} t_gpu {tg:.3f}s t_cpu {tc:.2f}s "
              f"acc_ms {st['ms_accumulate']:.2f} solve_ms {st['ms_target_solve']:.2f} pp {st['pixel_pulses']:.3g} "
              f"-> {st['pixel_pulses']/max(st['ms_accumulate'],1e-9)*1e3:.3g} pp/s", flush=True)

print(json.dumps(measure_peaks(0)), flush=True)
run("c1", pulses=768, bins=1024, out_lines=48, out_samples=160)
run("c1", pulses=2048, bins=4096, out_lines=64, out_samples=512)
run("c2", pulses=6144, bins=2048, out_lines=64, out_samples=512, n_targets=1)
run("c5", pulses=8192, bins=2048, out_lines=32, out_samples=256, n_targets=1, taps=8)
run("c5", pulses=8192, bins=2048, out_lines=32, out_samples=256, n_targets=1, taps=16)
run("c5", pulses=8192, bins=2048, out_lines=32, out_samples=256, n_targets=1, taps=32)
