"""GPU parity of the SPME reciprocal sum (row a21 of SURVEY.md section 8) against the oracle's
restatement of ewald_recip.f90, through crcl_ewald_recip: energy and gradient per image within 1e-10
relative; the oracle itself is pinned against the plain Ewald sum in tests/test_oracle_ewald.py."""
import numpy as np
import pytest

from tests import common as C

pytestmark = pytest.mark.gpu


def handle(gpu, E):
    g = gpu.RPMD(gpu.PES_NONE, 1, np.ones(1), 1.0, 1.0)
    g.set_ewald(dict(box=E.box, a_ewald=E.a_ewald, nfft=E.nfft, bsorder=E.bsorder, bsmod=E.bsmod))
    return g


@pytest.mark.parametrize("L,n,nimg", [((30.0, 30.0, 30.0), 60, 5), ((36.0, 33.0, 40.0), 200, 2), ((78.6, 78.6, 78.6), 3000, 3)])
def test_matches_oracle(gpu, oracle, L, n, nimg):
    E = oracle.Ewald(L)
    g = handle(gpu, E)
    rng = np.random.default_rng(n)
    x = rng.uniform(-0.3, 1.3, (nimg, n, 3)) * np.asarray(L)       # also atoms outside the box
    q = rng.normal(0, 0.4, n)
    q -= q.mean()
    e, grad = g.ewald_recip(x, q)
    for i in range(nimg):
        eo, go = E.recip(x[i], q)
        assert abs(e[i] - eo) < C.TOL_EG * abs(eo)
        assert np.abs(grad[i] - go).max() < C.TOL_EG * np.abs(go).max()


def test_spme_matches_plain_ewald_on_the_device(gpu, oracle):
    L = (30.0, 30.0, 30.0)
    E = oracle.Ewald(L)
    g = handle(gpu, E)
    rng = np.random.default_rng(3)
    x = rng.uniform(0, 1, (1, 60, 3)) * np.asarray(L)
    q = rng.normal(0, 0.4, 60)
    q -= q.mean()
    e, grad = g.ewald_recip(x, q)
    e2, g2 = E.direct_recip(x[0], q, 14)
    assert abs(e[0] - e2) < 1e-4 * abs(e2) and np.abs(grad[0] - g2).max() < 1e-2 * np.abs(g2).max()


def test_needs_setup_and_order_five(gpu):
    g = gpu.RPMD(gpu.PES_NONE, 1, np.ones(1), 1.0, 1.0)
    with pytest.raises(gpu.CaracalGpuError):
        g.ewald_recip(np.zeros((1, 2, 3)), np.array([1.0, -1.0]))
    with pytest.raises(gpu.CaracalGpuError):
        g.set_ewald(dict(box=[30.0] * 3, a_ewald=0.28, nfft=20, bsorder=4, bsmod=np.ones((3, 20))))
