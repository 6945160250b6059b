"""`src/simulator/disturbances.jl:4-84` restated on the host (`contactimplicitmpc.jl_b200/disturbances.py`): the index
bookkeeping of the open-loop disturbance (comment table of disturbances.jl:16-19), the impulse look-up, the random one."""
import numpy as np


def _mod():
    import importlib.util
    import os
    from common import ROOT
    spec = importlib.util.spec_from_file_location("cimpc_dist", os.path.join(ROOT, "contactimplicitmpc.jl_b200", "disturbances.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_open_loop_disturbance_indexing():
    d = _mod().OpenLoopDisturbance(np.array([[3.0], [6.0], [9.0], [12.0]]), 3)
    # With N_sample = 3: t = 1..10 → idx = 1 1 1 2 2 2 3 3 3 4 (disturbances.jl:16-19), value w[idx] / N_sample
    assert [float(d(t)[0]) for t in range(1, 11)] == [1, 1, 1, 2, 2, 2, 3, 3, 3, 4]
    assert float(d(1)[0]) == 1.0  # t == 1 resets


def test_impulse_and_random_disturbances():
    m = _mod()
    d = m.ImpulseDisturbance([[15.0, 25.0, 30.0], [0.0, 5.0, 0.0]], [20, 370])   # flat_trot.jl:64-75
    assert d(20).tolist() == [15.0, 25.0, 30.0] and d(370).tolist() == [0.0, 5.0, 0.0] and d(21).tolist() == [0.0, 0.0, 0.0]
    r = m.RandomDisturbance(2, [0.5], 100, 0.01, seed=3)
    w = np.array([r(t) for t in range(1, 50)])
    assert w.shape == (49, 2) and (w >= 0).all() and (w <= 0.5).all()
    assert np.array_equal(r(7), m.RandomDisturbance(2, [0.5], 100, 0.01, seed=3)(7))
    assert m.EmptyDisturbances(3)(5).tolist() == [0.0, 0.0, 0.0]
