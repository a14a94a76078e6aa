"""Test and bench tooling for the B200 TDBP backend -- NOT product code.

``synth``  deterministic synthetic scenes of the BASELINE.json configurations (orbit, radar
           grids, point-target echoes, DEMs) used by tests/, bench.py and smoke().
``irf``    point-target impulse-response metrics for the IRF parity gate.

Nothing under ``isce3_b200/`` imports this package.
"""
