#!/usr/bin/env python
"""The reference's OWN GPU path on this B200: the unmodified NREL/marbles sources compiled with nvcc for sm_100a
(oracle/refbuild `make cuda` -> oracle/_ref/marbles3d.cuda.ex; AMReX ParallelFor kernels, device arena), timed on
the same Taylor-Green deck as bench.py.  512^3 does not fit its ~190 words per cell in 180 GB, so it runs the
largest sizes that do.  Prints one JSON line per size.

    python profiles/ref_gpu.py [sizes ...]        (default 256 384)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

exe = os.path.join(ROOT, "oracle", "_ref", "marbles3d.cuda.ex")
sizes = [int(a) for a in sys.argv[1:]] or [256, 384]
for n in sizes:
    r = bench.reference_mlups(n, 10, threads=1, exe=exe)
    if r is None:
        print(json.dumps({"impl": "reference-cuda", "size": n, "unavailable": "executable missing or run failed"}))
        continue
    print(json.dumps({"impl": "reference-cuda", "exe": r[3], "size": n, "steps": 10, "value": r[0], "unit": "MLUPS",
                      "ms_per_step": r[1] * 1e3, "timing": "LBM::evolve() inclusive time of the reference's TinyProfiler"}))
