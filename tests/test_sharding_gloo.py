"""world_size-2 gloo test (CPU) of the N > 1 path of bench.py: the frame is cut into
contiguous azimuth blocks (bench.block_bounds + Scene.out_subgrid), each rank focuses its
block independently (here with the CPU oracle standing in for the device), results are
gathered with no data-path collective other than the final gather, timing is max-reduced."""
import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _worker(rank, world, store_path, result_path):
    sys.path.insert(0, str(ROOT))
    os.environ["OMP_NUM_THREADS"] = "2"
    import bench
    from testkit import synth
    from oracle import tdbp
    dist.init_process_group("gloo", init_method=f"file://{store_path}", rank=rank, world_size=world)
    sc = synth.make_scene("c2", pulses=512, bins=512, out_lines=11, out_samples=24, n_targets=1)
    lines = sc.out_geometry.grid_length
    a0, a1 = bench.block_bounds(lines, world, rank)
    sub = sc.out_subgrid(a0, a1)
    out = np.zeros((a1 - a0, sc.out_geometry.grid_width), np.complex64)
    tdbp.port().backproject(out, sub, sc.rc, sc.in_geometry, sc.dem, sc.fc, sc.ds, sc.kernel,
                            sc.dry_tropo_model)
    # gather blocks on rank 0 (the host-side gather of SURVEY.md 8e; no reduction of image data)
    blocks = [None] * world
    dist.all_gather_object(blocks, (a0, a1, out))
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # the max-over-ranks timing reduction of bench.py
    if rank == 0:
        full = np.zeros((lines, sc.out_geometry.grid_width), np.complex64)
        covered = np.zeros(lines, int)
        for b0, b1, blk in blocks:
            full[b0:b1] = blk
            covered[b0:b1] += 1
        np.savez(result_path, full=full, covered=covered, tmax=t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_azimuth_block_sharding_world_size_2(oracles):
    port, _ = oracles
    import bench
    from testkit import synth
    # block bounds tile the line range exactly, for ragged splits too
    for lines in (11, 8, 3, 1):
        for world in (1, 2, 4, 8):
            bounds = [bench.block_bounds(lines, world, r) for r in range(world)]
            assert bounds[0][0] == 0 and bounds[-1][1] == lines
            assert all(b[1] == c[0] for b, c in zip(bounds, bounds[1:]))
            assert max(b[1] - b[0] for b in bounds) - min(b[1] - b[0] for b in bounds) <= 1
    with tempfile.TemporaryDirectory() as d:
        store, res = os.path.join(d, "store"), os.path.join(d, "res.npz")
        mp.spawn(_worker, args=(2, store, res), nprocs=2, join=True)
        r = np.load(res)
    assert np.all(r["covered"] == 1) and r["tmax"][0] == 2.0
    sc = synth.make_scene("c2", pulses=512, bins=512, out_lines=11, out_samples=24, n_targets=1)
    whole = np.zeros((11, 24), np.complex64)
    port.backproject(whole, *sc.backproject_args())
    np.testing.assert_allclose(r["full"], whole, rtol=0, atol=1e-6 * np.abs(whole).max())
