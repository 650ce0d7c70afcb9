"""CPU: lidarFactor residuals + the Ceres trust-region restatement.  Nothing upstream pins these (the reference has no
tests and Ceres is absent), so the oracle carries self-consistency checks (SURVEY.md 8c i-iv)."""
import numpy as np
import pytest


def _blocks(orc, rng, n_edge=40, n_plane=60, n_pn=30, noise=0.05):
    out = []
    for _ in range(n_edge):
        cp = rng.uniform(-20, 20, 3); a = cp + rng.normal(0, noise, 3); b = a + rng.uniform(-1, 1, 3)
        out.append(orc.make_edge(cp, a, b))
    for _ in range(n_plane):
        cp = rng.uniform(-20, 20, 3); j = cp + rng.normal(0, noise, 3)
        out.append(orc.make_plane(cp, j, j + rng.uniform(-1, 1, 3), j + rng.uniform(-1, 1, 3)))
    for _ in range(n_pn):
        cp = rng.uniform(-20, 20, 3); n = rng.normal(0, 1, 3); n /= np.linalg.norm(n)
        out.append(orc.make_plane_norm(cp, n, -float(n @ cp) + rng.normal(0, noise)))
    return np.array(out)


def _x(rng, ang=0.05):
    d = rng.normal(0, ang, 3)
    q = np.concatenate([np.sin(np.linalg.norm(d)) * d / np.linalg.norm(d), [np.cos(np.linalg.norm(d))]])
    return np.concatenate([q, rng.normal(0, 0.3, 3)])


def test_autodiff_equals_closed_form(orc):
    rng = np.random.default_rng(1)
    bl = _blocks(orc, rng)
    for _ in range(3):
        x = _x(rng)
        r1, J1, c1 = orc.evaluate(bl, x, autodiff=True)    # Jet<7> through the literal functor text + manifold Jacobian
        r2, J2, c2 = orc.evaluate(bl, x, autodiff=False)   # closed form used by the CUDA kernel
        assert np.allclose(r1, r2, rtol=0, atol=1e-12) and abs(c1 - c2) < 1e-12 * max(c1, 1)
        assert np.abs(J1 - J2).max() < 1e-9


def test_jacobian_finite_differences(orc):
    rng = np.random.default_rng(2)
    bl = _blocks(orc, rng, 10, 10, 10)
    x = _x(rng)
    H = 1e9                                          # Huber knee far away => J is the plain derivative of r
    r0, J, _ = orc.evaluate(bl, x, huber=H)
    h = 1e-6
    for k in range(6):
        d = np.zeros(6); d[k] = h
        xp = np.concatenate([orc.quat_plus(x[:4], d[:3]), x[4:] + d[3:]])
        xm = np.concatenate([orc.quat_plus(x[:4], -d[:3]), x[4:] - d[3:]])
        fd = (orc.evaluate(bl, xp, huber=H)[0] - orc.evaluate(bl, xm, huber=H)[0]) / (2 * h)
        assert np.abs(fd - J[:, k]).max() < 1e-6 * max(1.0, np.abs(J).max())


def test_huber_and_normal_equations(orc):
    rng = np.random.default_rng(3)
    bl = _blocks(orc, rng, noise=0.3)                  # many residuals beyond the 0.1 knee
    x = _x(rng)
    r, J, cost = orc.evaluate(bl, x)
    JtJ, Jtr, c2 = orc.normal_equations(bl, x)
    assert np.allclose(JtJ, J.T @ J, rtol=1e-12) and np.allclose(Jtr, J.T @ r, rtol=1e-12) and cost == c2
    # cost = sum 0.5 rho(s) with rho Huber(0.1) per BLOCK (an edge's 3 rows are one block)
    raw = []
    i = 0
    x_r, _, _ = orc.evaluate(bl, x)
    for b in bl:
        rows = 3 if b[0] == orc.EDGE else 1
        one = np.array([b])
        rr, _, cc = orc.evaluate(one, x)
        raw.append(cc)
        i += rows
    assert abs(sum(raw) - cost) < 1e-12 * cost
    assert abs(orc.cost(bl, x) - cost) < 1e-13 * cost


def test_solve_monotone_and_matches_scipy(orc):
    from scipy.optimize import least_squares
    rng = np.random.default_rng(4)
    x_true = _x(rng, 0.02)
    # noise-free plane-norm + edge constraints generated from a known transform => exact minimum at x_true
    bl = []
    Rm = _rot(x_true[:4])
    for _ in range(80):
        cp = rng.uniform(-20, 20, 3); w = Rm @ cp + x_true[4:]
        n = rng.normal(0, 1, 3); n /= np.linalg.norm(n)
        bl.append(orc.make_plane_norm(cp, n, -float(n @ w)))
    for _ in range(40):
        cp = rng.uniform(-20, 20, 3); w = Rm @ cp + x_true[4:]
        d = rng.normal(0, 1, 3); d /= np.linalg.norm(d)
        bl.append(orc.make_edge(cp, w + 0.1 * d, w - 0.1 * d))
    bl = np.array(bl)
    x0 = np.array([0, 0, 0, 1.0, 0, 0, 0])
    x, s, trace = orc.solve(bl, x0, max_iters=50)
    costs = [row[0] for row in trace if row[7] == 1]
    assert all(b <= a for a, b in zip([trace[0][0]] + costs, costs)), "accepted steps never increase the cost"
    assert np.abs(x[4:] - x_true[4:]).max() < 1e-7 and abs(abs(x[:4] @ x_true[:4]) - 1) < 1e-12
    # 4-iteration budget of the reference (max_num_iterations = 4): terminates by iteration count
    x4, s4, t4 = orc.solve(bl, x0, max_iters=4)
    assert s4["num_iterations"] <= 4 and s4["final_cost"] < s4["initial_cost"]

    def fun(p):
        xx = np.concatenate([orc.quat_plus(x0[:4], p[:3]), p[3:]])
        return orc.evaluate(bl, xx)[0]
    sp = least_squares(fun, np.zeros(6), method="lm", xtol=1e-14, ftol=1e-14)
    xs = np.concatenate([orc.quat_plus(x0[:4], sp.x[:3]), sp.x[3:]])
    assert np.abs(xs[4:] - x[4:]).max() < 1e-6 and abs(abs(xs[:4] @ x[:4]) - 1) < 1e-10


def _rot(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def test_empty_problem_leaves_parameters(orc):
    x0 = np.array([0.1, 0, 0, 0.995, 1, 2, 3])
    x, s, _ = orc.solve(np.zeros((0, 11)), x0)
    assert np.array_equal(x, x0) and s["termination"] == 4


def test_golden_solve(orc):
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "solve_hdl64_pair.npz"))
    x, s, trace = orc.solve(g["blocks"], g["x0"], max_iters=4)
    assert np.allclose(x, g["x"], rtol=0, atol=1e-12)
    assert np.allclose(trace[:, [0, 5]], g["trace"][:, [0, 5]], rtol=1e-10)
    JtJ, Jtr, cost = orc.normal_equations(g["blocks"], g["x0"])
    assert np.allclose(JtJ, g["JtJ"], rtol=1e-12) and np.allclose(Jtr, g["Jtr"], rtol=1e-12)
