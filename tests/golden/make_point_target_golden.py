"""Generates tests/golden/point_target_golden.npz by running the REFERENCE's own
point-target analysis (python/packages/isce3/cal/point_target_info.py under
/root/reference) on a synthetic impulse response.  The reference module imports the
compiled isce3 extension for unrelated helpers; those imports are stubbed.  Run in the
build container only (the GPU box has no /root/reference); the .npz is committed."""
import importlib.util
import sys
import types
from pathlib import Path

import numpy as np

REF = Path("/root/reference/python/packages/isce3/cal/point_target_info.py")


def load_reference_module():
    for name in ("isce3", "isce3.core", "isce3.image", "isce3.image.v2", "isce3.io",
                 "isce3.io.dataset", "isce3.product"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["isce3.core"].DateTime = object
    sys.modules["isce3.core"].LUT2d = object
    sys.modules["isce3.image.v2"].resample_to_coords = None
    sys.modules["isce3.io.dataset"].DatasetReader = object
    sys.modules["isce3.product"].RadarGridParameters = object
    spec = importlib.util.spec_from_file_location("ref_point_target_info", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def synthetic_irf(n=160, i0=80.37, j0=79.81, bw_az=0.72, bw_rg=0.83, seed=7):
    i = np.arange(n)[:, None]
    j = np.arange(n)[None, :]
    z = np.sinc(bw_az * (i - i0)) * np.sinc(bw_rg * (j - j0))
    z = z * np.exp(1j * (0.31 * (i - i0) - 0.17 * (j - j0) + 0.4))
    rng = np.random.default_rng(seed)
    z = z + 1e-4 * (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    return z.astype(np.complex64)


def main():
    ref = load_reference_module()
    z = synthetic_irf()
    out = {"image": z}
    for tag, kw in (("a", dict(nov=32, chipsize=64)), ("b", dict(nov=16, chipsize=32, predict_null=True, fs_bw_ratio=1.2))):
        info, _ = ref.analyze_point_target(z, 80, 80, **kw)
        vals = []
        for axis in ("azimuth", "range"):
            for key in ("index", "offset", "resolution", "PSLR", "ISLR", "phase ramp"):
                vals.append(float(info[axis][key]))
        vals += [float(info["magnitude"]), float(info["phase"])]
        out[f"info_{tag}"] = np.array(vals)
    out["oversampled_4_decimated"] = ref.oversample(z[48:112, 48:112].copy(), 4)[::8, ::8]
    np.savez_compressed(Path(__file__).with_name("point_target_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
