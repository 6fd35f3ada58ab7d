"""CPU suite: the EMD oracle (oracle/emd_numpy.py) against known answers -- the auction algorithm with a small eps and
enough iterations reaches the OPTIMAL assignment (Bertsekas: within n * eps of it), which scipy's Hungarian solver
gives independently; and the structural properties of emd_linear/emd_cuda.cu the B200 kernel relies on."""
import numpy as np
from scipy.optimize import linear_sum_assignment

from oracle import emd_numpy as E


def test_auction_reaches_the_optimal_assignment_on_small_clouds():
    rng = np.random.default_rng(3)
    for n in (6, 17, 40):
        a, b = rng.random((n, 3)).astype(np.float32), rng.random((n, 3)).astype(np.float32)
        dist, asg = E.emd_auction(a, b, 1e-5, 20000)
        assert sorted(asg.tolist()) == list(range(n)), "a bijection once every source is assigned"
        cost = np.sqrt(((a[:, None, :] - b[None, :, :]) ** 2).sum(-1))
        r, c = linear_sum_assignment(cost)
        assert abs(float(np.sqrt(dist).sum()) - float(cost[r, c].sum())) <= n * 1e-5 + 1e-5
        assert np.allclose(dist, ((a - b[asg]) ** 2).sum(-1), rtol=1e-6, atol=1e-7)


def test_last_iteration_assigns_every_remaining_source_to_its_bid():
    rng = np.random.default_rng(4)
    a, b = rng.random((64, 3)).astype(np.float32), rng.random((64, 3)).astype(np.float32)
    dist, asg = E.emd_auction(a, b, 0.005, 1)
    # one iteration = the last one: everybody takes the nearest target at zero prices (emd_cuda.cu:196-199)
    nn = ((a[:, None, :] - b[None, :, :]) ** 2).sum(-1).argmin(1)
    assert np.array_equal(asg, nn)
    d3, a3 = E.emd_auction(a, b, 0.005, 3)
    assert (a3 >= 0).all() and len(np.unique(a3)) >= len(np.unique(asg))


def test_gradient_restatement():
    rng = np.random.default_rng(5)
    a, b = rng.random((32, 3)).astype(np.float32), rng.random((32, 3)).astype(np.float32)
    dist, asg = E.emd_auction(a, b, 0.01, 50)
    g = rng.standard_normal(32).astype(np.float32)
    num = np.zeros_like(a)
    h = 1e-3
    for j in range(32):
        for c in range(3):
            ap = a.copy(); ap[j, c] += h
            am = a.copy(); am[j, c] -= h
            num[j, c] = g[j] * (((ap[j] - b[asg[j]]) ** 2).sum() - ((am[j] - b[asg[j]]) ** 2).sum()) / (2 * h)
    assert np.allclose(E.emd_grad(a, b, g, asg), num, rtol=1e-2, atol=1e-3)
