"""Host side of the rate pipeline (caracal_b200/rate.py): window grid,
umbrella integration, extremum search and k(T) formula on closed-form inputs, and the complete
calc_rate sequence on the oracle behind the product's host interface (tests/oracle_handle.py)."""
import math

import numpy as np

from caracal_b200 import rate as R
from tests import common as C


def test_window_grid_of_the_shipped_example():
    # examples/calc_rate/h+h2/rate.key: bonds -0.05 1.05, dist 0.01
    n_over, n_samplings, n_all, xi = R.window_grid(-0.05, 1.05, 0.01)
    assert (n_over, n_samplings, n_all) == (5, 105, 111)
    assert len(xi) == 110 and abs(xi[0] + 0.05) < 1e-12 and abs(xi[-1] - 1.04) < 1e-12
    assert np.allclose(np.diff(xi), 0.01)
    assert abs(xi[105] - 1.0) < 1e-12          # first "over" window sits at the TS


def test_umbrella_integration_recovers_a_known_pmf():
    # W(xi) = h sin^2(pi xi / 2) on [0,1]; in a stiff harmonic window the biased distribution is a
    # normal with mean xi0 - W'(xi0)/k (first order) and variance 1/(beta (k + W''))
    beta, h, k = C.beta_calc_rate(300.0), 0.02, 15.0
    n_over, n_samplings, n_all, xi = R.window_grid(-0.05, 1.05, 0.01)
    W1 = lambda x: h * math.pi / 2 * np.sin(math.pi * x)
    W2 = lambda x: h * (math.pi ** 2) / 2 * np.cos(math.pi * x)
    mean = xi.copy()
    for _ in range(50):                                   # solve W'(m) + k (m - xi0) = 0
        mean = xi - W1(mean) / k
    var = 1.0 / (beta * (k + W2(mean)))
    bc, pmf = R.umbrella_integration(xi, mean, var, k, beta, -0.05, 1.05, 2000, 10, 20000)
    maxloc, minloc, xb = R.locate_extrema(bc, pmf, -0.05, 1.05, "ZERO")
    assert abs(xb - 1.0) < 5e-3
    assert abs(bc[minloc]) < 2e-3                         # 'ZERO': the bin at xi = 0
    assert abs((pmf[maxloc] - pmf[minloc]) - h) < 2e-4
    assert pmf.min() == 0.0


def test_calc_k_t_formula():
    beta = C.beta_calc_rate(300.0)
    m = C.masses("h3")
    k_t, k_mol = R.calc_k_t(0.767, 48.5 / R.HARTREE_KJ, 0.0, beta, [m[0] + m[1], m[2]], 16.0, 2)
    my_R = (m[0] + m[1]) * m[2] / (m[0] + m[1] + m[2])
    au = 2 * 0.767 * 4 * math.pi * 256.0 * math.sqrt(1 / (2 * math.pi * beta * my_R)) * math.exp(-beta * 48.5 / R.HARTREE_KJ)
    assert abs(k_mol / (au * 1e6 * 5.2917721092e-11 ** 3 / 2.418884326505e-17) - 1.0) < 1e-6
    assert 1e-17 < k_mol < 1e-14                          # H + H2 at 300 K: ~1e-16 cm^3/(molecule s)


def test_calc_k_t_unimolecular_formula():
    # Eyring: k_B T / h = 6.25e12 1/s at 300 K; a 50 kJ/mol barrier gives exp(-20.05)
    k = R.calc_k_t_unimol(1.0, 50.0 / 2625.50, 0.0, 300.0, 1)
    assert abs(k / (6.2509e12 * math.exp(-50.0 / (0.00831447 * 300.0))) - 1.0) < 1e-4


def test_whole_pipeline_on_the_oracle(oracle):
    from tests.oracle_handle import OracleRPMD
    name, nb, kelvin = "h3", 4, 300.0
    m, mech = C.masses(name), C.mechanism(name)
    beta, dt = C.beta_calc_rate(kelvin), C.dt_au(0.1)
    g, g1 = OracleRPMD(name, nb, m, beta, dt), OracleRPMD(name, 1, m, beta, dt)
    for h in (g, g1):
        h.set_mechanism(mech)
        h.set_seed(C.SEED)
    out = R.calc_rate(g, g1, C.h3_ts(), m, mech, kelvin, beta, umbr_lo=0.9, umbr_hi=1.02, umbr_dist=0.01,
                      gen_steps=40, equi_steps=20, umbr_steps=60, umbr_traj=2, xi_min=0.9, xi_max=1.02, nbins=200,
                      recr_equi=30, child_tot=16, child_interv=9, child_point=4, child_evol=30, andersen_step=10,
                      npaths=2, pmf_minloc="PMF_MIN")
    assert len(out["xi_wins"]) == 12 and out["struc_equi"].shape == (12, 3, 3)
    assert np.abs(out["start_xis"] - out["xi_wins"]).max() < 0.05      # the bias holds the windows
    assert np.abs(out["average"] - out["xi_wins"]).max() < 0.05
    assert (out["variance"] > 0).all() and (out["variance"] < 1e-2).all()
    assert np.abs((out["struc_equi"] * m[None, :, None]).sum(axis=1)).max() < 1e-9   # COM removed
    assert out["kappa_t"].shape == (30,) and abs(out["kappa_t"][0] - 1.0) < 0.2
    assert 0.0 < out["kappa"] <= 1.2 and out["k_t_molec"] > 0
