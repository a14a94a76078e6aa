#!/usr/bin/env python
"""Full-size parity check of the BASELINE.json configurations on a B200 (not part of the
quick test suite: minutes of CPU oracle time).  For each configuration the GPU focuses the
whole frame through ``backproject()``; the CPU oracle (oracle/_ref when present) focuses

  * a chip around EVERY point target  -> peak phase, IRF peak location, PSLR, ISLR (both axes)
  * a few full-width azimuth lines    -> NaN masks, relative RMS error, height layer

Gate (BASELINE.json): rel. RMS <= 1e-4, peak phase <= 1 mrad, peak location <= 0.01 sample,
PSLR / ISLR <= 0.05 dB.  Writes one JSON line per configuration and a markdown table.

    python scripts/validate_configs.py [c1 c2 c4 c5k8 c5k16 c5k32] [--out profiles/r01_config_parity]
"""
import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from testkit.validate import CONFIGS, run  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=list(CONFIGS))
    ap.add_argument("--out", default="gpurun_out/config_parity")
    args = ap.parse_args()
    recs = [run(c) for c in args.configs]
    Path(args.out + ".jsonl").write_text("".join(json.dumps(r) + "\n" for r in recs))
    md = ["| config | frame | taps | DEM | kernel | pixel·pulses | GPU pp/s (one-shot) | targets | rel RMS (block) | worst chip rel RMS | peak phase (rad) | peak shift az / rg | ΔPSLR dB | ΔISLR dB | Δheight m | pass |",
          "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    for r in recs:
        md.append(f"| {r['config']} | {r['pulses']}×{r['bins']} → {r['out_lines']}×{r['out_samples']} | {r['taps']} | {r['dem']} | "
                  f"{'fast' if r['used_fast_kernel'] else 'generic'}{'/imm' if r['fast_variant'] >= 0 else ''} | {r['pixel_pulses']:.3g} | "
                  f"{r['gpu_pp_per_s']:.3g} | {r['targets_checked']} | {r['rel_rms_block']:.2e} | {r['worst_chip_rel_rms']:.2e} | "
                  f"{r['worst_peak_phase_rad']:.1e} | {r['worst_az_peak_offset_diff']:.3f} / {r['worst_rg_peak_offset_diff']:.3f} | "
                  f"{r['worst_pslr_diff_db']:.4f} | {r['worst_islr_diff_db']:.4f} | {r['height_max_abs_diff_m']:.1e} | {'yes' if r['pass'] else 'NO'} |")
    Path(args.out + ".md").write_text("\n".join(md) + "\n")
    print("\n".join(md))
    return 0 if all(r["pass"] for r in recs) else 1


if __name__ == "__main__":
    sys.exit(main())
