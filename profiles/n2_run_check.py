"""2-rank `python -m marbles_b200.run` against the single-rank run of the same deck: plotfiles agree (run under torchrun)"""
import os, sys, subprocess, numpy as np, tempfile
sys.path.insert(0, os.getcwd())
from oracle import oracle as O
z = np.load("tests/golden/tg12.npz")
work = tempfile.mkdtemp()
open(work + "/deck.inp", "w").write(str(z["deck"]))
ov = ["amr.n_cell=32 32 48", "max_step=6", "amr.plot_int=3", "amr.chk_int=-1", "amr.max_grid_size=16"]
env = dict(os.environ, PYTHONPATH=os.getcwd())
for tag, cmd in (("one", [sys.executable, "-m", "marbles_b200.run"]),
                 ("two", [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29655", "-m", "marbles_b200.run"])):
    d = work + "/" + tag
    os.makedirs(d)
    r = subprocess.run(cmd + [work + "/deck.inp"] + ov, cwd=d, env=env, capture_output=True, text=True)
    print(tag, "rc", r.returncode, sorted(os.listdir(d)), r.stderr[-300:] if r.returncode else "")
a, b = O.read_plotfile(work + "/one/plt00006"), O.read_plotfile(work + "/two/plt00006")
worst = max(np.abs(a[k] - b[k]).max() / max(np.abs(a[k]).max(), 1e-30) for k in a["__names__"] if not k.startswith("vort"))
print("fields 1 rank vs 2 ranks, worst relative difference (vorticity excluded):", worst, "files", sorted(os.listdir(work + "/two/plt00006/Level_0")))
