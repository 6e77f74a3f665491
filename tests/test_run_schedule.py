"""Host logic of marbles_b200.run.evolve on CPU: the cadence of plotfiles / checkpoints / forces lines against a
literal per-step transcription of the reference's init_data + evolve (Source/LBM.cpp:155-195, 398-448), with a
stand-in for the GPU object (no device work)."""
import itertools
import os

import pytest

from marbles_b200 import run as R


class FakeInputs:
    def __init__(self, deck):
        self.deck = deck


class FakeLBM:
    """counts steps; records every call that the driver makes"""

    def __init__(self, deck):
        self.inp, self.world, self.isteps, self.time, self.dt, self.log = FakeInputs(deck), 1, 0, 0.0, 1.0, []

    def init_data(self):
        self.isteps, self.time = 0, 0.0

    def read_checkpoint_file(self, path):
        self.isteps = int(os.path.basename(path)[-5:])
        self.time = float(self.isteps)

    def f_to_macrodata(self):
        pass

    def step(self, n, want_macrodata=False):
        assert n >= 1
        self.isteps += n
        self.time += n * self.dt
        self.log.append(("step", n, bool(want_macrodata)))

    def write_checkpoint_file(self, out_dir, prefix, digits=5):
        return os.path.join(out_dir, f"{prefix}{self.isteps:0{digits}d}")

    def compute_eb_forces(self):
        return [0.0, 0.0, 0.0]


def reference_schedule(max_step, plot_int, chk_int, stop_time, restart_step):
    """what the reference writes, in order (names only)"""
    out, step0 = [], restart_step or 0
    if restart_step is None and chk_int > 0:
        out.append("chk%05d" % 0)
    if plot_int > 0:
        out.append("plt%05d" % step0)
    isteps, t, last_plot = step0, float(step0), 0
    step = isteps
    while step < max_step and t < stop_time:
        isteps += 1
        t += 1.0
        if plot_int > 0 and (step + 1) % plot_int == 0:
            last_plot = step + 1
            out.append("plt%05d" % isteps)
        if chk_int > 0 and (step + 1) % chk_int == 0:
            out.append("chk%05d" % isteps)
        if t >= stop_time - 1.0e-6:
            break
        step += 1
    if plot_int > 0 and isteps > last_plot:
        out.append("plt%05d" % isteps)
    return out, isteps


@pytest.mark.parametrize("max_step,plot_int,chk_int,stop_time,restart", [
    (10, 5, 5, None, None), (10, 4, 3, None, None), (7, 10, -1, None, None), (12, -1, 5, None, None), (9, 3, 2, 6.0, None),
    (10, 5, 5, None, 5), (11, 4, -1, None, 8), (1, 1, 1, None, None), (6, 2, 3, 100.0, None), (0, 2, 2, None, None),
    (20, 4, -1, 6.0, None), (20, 5, 3, 7.0, None), (9, 4, -1, 3.0, None), (30, 7, -1, 20.0, 14)])
def test_output_cadence_matches_reference_loop(tmp_path, monkeypatch, max_step, plot_int, chk_int, stop_time, restart):
    deck = {"max_step": [str(max_step)], "amr.plot_int": [str(plot_int)], "amr.chk_int": [str(chk_int)]}
    if stop_time is not None:
        deck["stop_time"] = [str(stop_time)]
    if restart is not None:
        deck["amr.restart"] = ["chk%05d" % restart]
    lbm = FakeLBM(deck)
    monkeypatch.setattr(R, "write_lbm_plotfile", lambda l, d, p, digits=5: os.path.join(d, f"{p}{l.isteps:05d}"))
    written = [os.path.basename(p) for p in R.evolve(lbm, str(tmp_path), log=lambda s: None)]
    want, final = reference_schedule(max_step, plot_int, chk_int, stop_time if stop_time is not None else float("inf"), restart)
    assert written == want and lbm.isteps == final
    # macrodata is requested for every step that ends in a plotfile: on the plot cadence, and for the closing
    # plotfile when max_step or stop_time end the run between two plot steps (the reference recomputes macrodata
    # every step, so its last plotfile never holds an older step's rho / vel / T)
    done = restart or 0
    plotted = {int(w[3:]) for w in want if w.startswith("plt")}
    for _, n, macro in lbm.log:
        done += n
        if done in plotted:
            assert macro, (done, lbm.log)


def test_forces_lines_one_per_step(tmp_path, monkeypatch):
    deck = {"max_step": ["5"], "amr.plot_int": ["-1"], "amr.chk_int": ["-1"], "lbm.compute_forces": ["1"]}
    lbm = FakeLBM(deck)
    R.evolve(lbm, str(tmp_path), log=lambda s: None)
    lines = open(tmp_path / "forces.txt").read().split("\n")
    assert len(lines) == 1 + 6 + 1 and all(len(l) == 96 for l in lines[:-1])
    assert [n for _, n, _ in lbm.log] == [1] * 5
    assert lines[0].split() == ["time", "fx", "fy", "fz"] and lines[3].split()[0] == "2"
    _ = itertools


def test_stop_time_between_lattice_steps_is_refused(tmp_path):
    """the reference shortens its last step to dt = stop_time - t (compute_dt); whole lattice steps only here"""
    from marbles_b200.lbm import MarblesError
    lbm = FakeLBM({"max_step": ["20"], "amr.plot_int": ["4"], "stop_time": ["6.5"]})
    with pytest.raises(MarblesError, match="whole number"):
        R.evolve(lbm, str(tmp_path), log=lambda s: None)


def test_file_name_digits(tmp_path, monkeypatch):
    deck = {"max_step": ["4"], "amr.plot_int": ["2"], "amr.chk_int": ["4"], "amr.file_name_digits": ["7"]}
    lbm = FakeLBM(deck)
    from marbles_b200 import plotfile as P
    monkeypatch.setattr(R, "write_lbm_plotfile", lambda l, d, p, digits=5: os.path.join(d, P.plot_file_name(p, l.isteps, digits)))
    lbm.write_checkpoint_file = lambda d, p, digits=5: os.path.join(d, P.chk_file_name(p, lbm.isteps, digits))
    written = [os.path.basename(p) for p in R.evolve(lbm, str(tmp_path), log=lambda s: None)]
    assert written == ["chk0000000", "plt0000000", "plt0000002", "plt0000004", "chk0000004"]
