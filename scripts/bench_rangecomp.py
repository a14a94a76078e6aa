"""Developer tool: throughput of the GPU RangeComp on C2-sized input (16384 lines x 12288
samples, 2049-sample chirp, mode Valid), device time by CUDA events around the kernels+FFTs,
and the numpy oracle timed on a few lines for scale."""
import json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from isce3_b200.focus import RangeComp, form_linear_chirp
from oracle import rangecomp as orc

lines, n, batch = 16384, 12288, 2048
fs, bw, dur = 24e6, 20e6, 85e-6
chirp = form_linear_chirp(bw / dur, dur, fs)
rng = np.random.default_rng(1)
x = (rng.standard_normal((batch, n), dtype=np.float32) + 1j * rng.standard_normal((batch, n), dtype=np.float32)).astype(np.complex64)
rc = RangeComp(chirp, n, maxbatch=batch, mode=RangeComp.Mode.Valid)
y = np.zeros((batch, rc.output_size), np.complex64)
for _ in range(2):
    rc.rangecompress(y, x)
ms, wall = [], []
for _ in range(lines // batch):
    t = time.perf_counter(); rc.rangecompress(y, x); wall.append(time.perf_counter() - t); ms.append(rc.last_device_ms())
t = time.perf_counter(); want = orc.rangecompress(chirp, x[:64], orc.VALID); t_cpu = time.perf_counter() - t
rel = float(np.linalg.norm(y[:64] - want) / np.linalg.norm(want))
dev_s, wall_s = sum(ms) * 1e-3, sum(wall)
alg_bytes = lines * 8.0 * (n + rc.output_size)
print(json.dumps({"workload": f"{lines} lines x {n} samples, chirp {chirp.size}, nfft {rc.fft_size}, mode valid, batch {batch}",
                  "device_ms_total": sum(ms), "lines_per_s_device": lines / dev_s, "lines_per_s_host_buffers": lines / wall_s,
                  "algorithmic_GBps_device": alg_bytes / dev_s / 1e9, "numpy_oracle_lines_per_s_1core": 64 / t_cpu,
                  "rel_rms_vs_oracle": rel}))
