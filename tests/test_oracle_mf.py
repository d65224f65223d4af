"""The matrix-free oracle (orc_apply_mf, BASELINE config E) on CPU: its three summation orders -- one per device kernel
generation: 0 k_apply_mf, 1 k_apply_mf2, 2 k_apply_mf3 / k_apply_mf4 (FP64 tensor cores) -- are roundings of ONE operator,
and that operator is MATMUL(storkm_pp, pmul) of p121.f90:93-97 (the stored-path oracle, pinned to the reference's goldens in
test_oracle_golden.py).  This is what anchors the matrix-free mirror to the reference instead of to the kernel alone."""
import numpy as np
import pytest

import oracle

E, V = 100.0, 0.3


def _coords(nod, distort):
    m = oracle.cube_p121(3, 2, 4, nod, aa=1.0, bb=2.0, cc=0.5)
    g = np.array(m.g_coord_pp, dtype=np.float64)           # (nels, 3, nod) = g_coord_pp(nod,3,iel)
    if distort:
        # the same displacement for every copy of a node, so the mesh stays conforming
        rng = np.random.RandomState(7)
        shift = distort * 0.25 * rng.uniform(-1, 1, (m.nn + 1, 3))
        g = g + np.transpose(shift[m.g_num_pp], (0, 2, 1))
    return g


@pytest.mark.parametrize("nod", [8, 20])
@pytest.mark.parametrize("distort", [0.0, 0.2])
def test_orders_are_roundings_of_the_stored_operator(nod, distort):
    g = _coords(nod, distort)
    nels, ntot = g.shape[0], 3 * nod
    g = np.ascontiguousarray(g)
    pm = np.random.RandomState(3).randn(nels, ntot)
    km = oracle.form_km_elastic(g, nod, 8, E, V)
    stored = oracle.matvec(km, pm)
    scale = np.abs(stored).max()
    outs = []
    for order in (0, 1, 2):
        oracle.set_mf_order(order)
        out = np.empty_like(pm)
        rc = oracle.lib().orc_apply_mf(nels, nod, 8, oracle._p(oracle._f64(g)), E, V, oracle._p(oracle._f64(pm)), oracle._p(out))
        assert rc == 0
        assert np.abs(out - stored).max() <= 2e-13 * scale, order
        outs.append(out)
    oracle.set_mf_order(2)
    # the orders differ (they are different fma chains) but only in the last bits
    assert np.abs(outs[2] - outs[0]).max() <= 1e-13 * scale
    assert np.abs(outs[2] - outs[1]).max() <= 1e-13 * scale


def test_default_order_follows_the_kernel_selection(monkeypatch):
    monkeypatch.delenv("PF_MF", raising=False)
    assert oracle.default_mf_order(20) == 2 and oracle.default_mf_order(8) == 2
    monkeypatch.setenv("PF_MF", "3")
    assert oracle.default_mf_order(20) == 2
    monkeypatch.setenv("PF_MF", "2lane")
    assert oracle.default_mf_order(20) == 1 and oracle.default_mf_order(8) == 0
    monkeypatch.setenv("PF_MF", "1lane")
    assert oracle.default_mf_order(20) == 0
