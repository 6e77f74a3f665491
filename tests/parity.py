"""Shared helpers of the parity tests: field scales and the stated tolerance.

Tolerance (BASELINE.json north_star): fp64 agreement <= 1e-12 per step on macrodata, relative to
the field's scale.  A field that is ~0 by symmetry (v in a 1-D tube, QCorr at T ~ 1/3) carries O(1)
RELATIVE round-off noise in any two summation orders (hard part 6), so the
difference is measured against max(|field|_inf, natural scale of its group)."""
import numpy as np

TOL_PER_STEP = 1.0e-12


def scales(fields: dict, R: float, gamma: float, inv_dx: float = 1.0) -> dict:
    rho0 = float(np.abs(fields["rho"]).max())
    tmax = float(np.abs(fields["temperature"]).max())
    cs = float(np.sqrt(gamma * R * tmax))
    e0 = float(np.abs(fields["two_rho_e"]).max())
    fmax = max(float(np.abs(fields[f"f_{q:02d}"]).max()) for q in range(27))
    gmax = max(float(np.abs(fields[f"g_{q:02d}"]).max()) for q in range(27))
    sc = {"rho": rho0, "two_rho_e": e0, "temperature": tmax}
    for k in ("vel_x", "vel_y", "vel_z", "vel_mag"):
        sc[k] = cs
    for k in ("QCorrX", "QCorrY", "QCorrZ"):
        sc[k] = rho0 * cs
    for k in ("dQCorrX", "dQCorrY", "dQCorrZ"):
        sc[k] = rho0 * cs * inv_dx
    for k in ("pxx", "pyy", "pzz", "pxy", "pxz", "pyz"):
        sc[k] = rho0
    for k in ("qx", "qy", "qz"):
        sc[k] = e0 * cs
    for k in ("vort_x", "vort_y", "vort_z", "vort_mag"):
        sc[k] = cs * inv_dx
    for q in range(27):
        sc[f"f_{q:02d}"] = fmax
        sc[f"g_{q:02d}"] = gmax
    return sc


def compare(mine: dict, ref: dict, sc: dict, nsteps: int, keys=None, tol_per_step: float = TOL_PER_STEP):
    """returns (worst relative error, its field); asserts the tolerance"""
    worst, wkey = 0.0, None
    for k in (keys or mine.keys()):
        if k not in ref or k not in sc:
            continue
        denom = max(float(np.abs(ref[k]).max()), sc[k])
        err = float(np.abs(np.asarray(mine[k]) - np.asarray(ref[k])).max()) / denom
        if err > worst:
            worst, wkey = err, k
    tol = tol_per_step * max(nsteps, 1)
    assert worst <= tol, f"field {wkey}: error {worst:.3e} of its scale after {nsteps} steps (tolerance {tol:.1e})"
    return worst, wkey
