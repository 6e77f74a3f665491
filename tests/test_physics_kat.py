"""Physics known-answer tests (the reference's Tools/viscosity_test and Tools/thermaldiffusivity_test notebooks,
SURVEY section 8c iii): a shear wave decays as exp(-nu k^2 t) and a temperature wave at constant pressure as
exp(-alpha k^2 t).  The decay rate is fitted to the first Fourier mode.  On CPU this pins the oracle (the
reference's arithmetic); on the GPU the CUDA path must give the same rate as the oracle and the nominal
coefficient within the discretisation error of a 20 / 24-cell wave."""
import numpy as np
import pytest

from conftest import load_golden

STEPS = list(range(0, 401, 50))
CASES = {  # golden deck, overrides, field, nominal coefficient, allowed relative deviation from nominal
    "viscosity": ("shear", ["lbm.nu = 0.02"], "vel_x", 0.02, 0.03),
    "thermal_diffusivity": ("thermal", ["lbm.alpha = 0.01", "lbm.nu = 0.01"], "temperature", 0.01, 0.02),
}


def fitted_coefficient(stepper, fields, field):
    amp, done = [], 0
    for s in STEPS:
        if s > done:
            stepper(s - done)
            done = s
        line = fields()[field][0, :, 0]
        amp.append(np.abs(np.fft.rfft(line)[1]) * 2 / len(line))
    k = 2 * np.pi / len(line)
    return -np.polyfit(STEPS, np.log(np.array(amp)), 1)[0] / k ** 2


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_transport_coefficient(oracle_mod, name):
    O = oracle_mod
    case, ov, field, nominal, tol = CASES[name]
    _, deck_text, _ = load_golden(case)
    o = O.Oracle(O.lbm_setup(O.parse_deck(None, deck_text.splitlines() + ov)))
    o.initialize()
    got = fitted_coefficient(o.step, o.fields, field)
    assert abs(got - nominal) <= tol * nominal, (name, got)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_gpu_transport_coefficient(oracle_mod, name):
    from marbles_b200.inputs import parse_deck
    from marbles_b200.lbm import LBM
    O = oracle_mod
    case, ov, field, nominal, tol = CASES[name]
    _, deck_text, _ = load_golden(case)
    o = O.Oracle(O.lbm_setup(O.parse_deck(None, deck_text.splitlines() + ov)))
    o.initialize()
    ref = fitted_coefficient(o.step, o.fields, field)
    lbm = LBM(parse_deck(text=deck_text, overrides=ov))
    lbm.init_data()
    def fields():
        if lbm.isteps == 0:
            lbm.f_to_macrodata()  # macrodata of the initial state
        return lbm.fields()

    got = fitted_coefficient(lambda n: lbm.step(n, want_macrodata=True), fields, field)
    lbm.close()
    print(f"{name}: CUDA {got:.8g}, oracle {ref:.8g}, nominal {nominal}")
    assert abs(got - nominal) <= tol * nominal
    assert abs(got - ref) <= 1e-7 * nominal
