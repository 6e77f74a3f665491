"""`python -m marbles_b200.fcompare pltA pltB [--rel-tol R] [--abs-tol A]` -- the comparison AMReX's
`fcompare` (Tools/Plotfile) makes between two plotfiles, for the single-level files this package and the reference
write: per variable the largest absolute difference and that difference relative to the largest magnitude in
the first file; exit status 0 if every variable is within the tolerances (default: identical)."""
from __future__ import annotations

import argparse
import sys

import numpy as np

from .plotfile import read_plotfile


def compare(path_a: str, path_b: str):
    a, b = read_plotfile(path_a), read_plotfile(path_b)
    if a["__names__"] != b["__names__"]:
        raise ValueError("the plotfiles hold different variables")
    rows = []
    for name in a["__names__"]:
        if a[name].shape != b[name].shape:
            raise ValueError(f"{name}: grids differ, {a[name].shape} vs {b[name].shape}")
        d = float(np.abs(a[name] - b[name]).max())
        mag = float(np.abs(a[name]).max())
        rows.append((name, d, d / mag if mag > 0 else (0.0 if d == 0 else float("inf"))))
    return rows, (a["__time__"], b["__time__"])


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("plotfile_a")
    ap.add_argument("plotfile_b")
    ap.add_argument("--rel-tol", type=float, default=0.0)
    ap.add_argument("--abs-tol", type=float, default=0.0)
    args = ap.parse_args(argv)
    rows, times = compare(args.plotfile_a, args.plotfile_b)
    print(f"{'variable name':24s} {'absolute error':>24s} {'relative error':>24s}")
    print(f"{'':24s} {'(||A - B||)':>24s} {'(||A - B||/||A||)':>24s}")
    print("-" * 74)
    print(" level = 0")
    bad = 0
    for name, d, r in rows:
        print(f" {name:23s} {d:24.16g} {r:24.16g}")
        if d > args.abs_tol and r > args.rel_tol:
            bad += 1
    if times[0] != times[1]:
        print(f"times differ: {times[0]} vs {times[1]}")
        bad += 1
    print("PLOTFILES AGREE" if bad == 0 else f"PLOTFILES DISAGREE in {bad} place(s)")
    return 0 if bad == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
