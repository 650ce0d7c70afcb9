"""include/lidarFactor.hpp (the boundary's second preserved surface, reference src/lidarFactor.hpp:12-138): compiles without
Ceres / Eigen, its templated operator() == the oracle's functors, its analytic Evaluate() == Jet autodiff through the
quaternion manifold -- also for the interpolation ratio s != 1 of the DISTORTION build -- and (GPU) the blocks it packs give
the same normal equations through aloam_normal_equations."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cases(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("lf") / "lidar_factor_check")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle"),
                           "-o", exe, os.path.join(ROOT, "tests", "cpp", "lidar_factor_check.cc")])
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.splitlines()
    recs, x = [], None
    it = iter(out)
    summary = None
    for line in it:
        if line.startswith("X "):
            x = np.array(line.split()[1:], float)
        elif line.startswith("B "):
            b = np.array(line.split()[1:], float)
            r = np.array(next(it).split()[1:], float)
            j = np.array(next(it).split()[1:], float).reshape(-1, 6)
            recs.append((x, b, r, j))
        elif line.startswith("WORST"):
            p = line.split()
            summary = (float(p[1]), float(p[3]), int(p[5]))
    return recs, summary


def test_header_compiles_without_ceres_and_matches_autodiff(cases):
    recs, (worst, worst_tiny, n) = cases
    assert n == 900 and len(recs) == 900
    assert worst < 1e-10          # analytic Jacobian vs Jet<7> autodiff x manifold Jacobian, residuals identical
    assert worst_tiny < 1e-6      # 1e-7 rad rotations with s != 1: the autodiff side is ill-conditioned (acos near 1)


def test_header_functors_equal_oracle_functors(cases, orc):
    """residuals and tangent Jacobians of the header == the oracle's restatement of the reference functors (autodiff)"""
    recs, _ = cases
    worst = 0.0
    for x, b, r, j in recs[::7]:
        # rotations of ~1e-7 rad with s != 1: Eigen's slerp switches to a linear blend / acos loses digits -- the autodiff side is only good to 1e-8 there
        tiny = 0.0 < np.linalg.norm(x[:3]) < 1e-6 and b[0] != 2 and b[10] != 1.0
        ro, jo, _ = orc.evaluate(b.reshape(1, -1), x, huber=1e12, autodiff=True)
        assert np.abs(ro - r).max() < 1e-12
        if not tiny:
            worst = max(worst, float(np.abs(jo - j).max()))
    assert worst < 1e-10


def test_header_constructors_accept_any_xyz_vector():
    """the constructors take anything with x() y() z() (Eigen::Vector3d in the reference call sites, laserOdometry.cpp:373-381)"""
    src = r'''
#include "lidarFactor.hpp"
struct MyVec { double a, b, c; double x() const { return a; } double y() const { return b; } double z() const { return c; } };
int main() {
  MyVec p{1, 2, 3}, a{1, 2, 4}, b{1, 3, 4}, m{2, 2, 4};
  LidarEdgeFactor e(p, a, b, 1.0);
  LidarPlaneFactor f(p, a, b, m, 1.0);
  LidarPlaneNormFactor g(p, MyVec{0, 0, 1}, -3.0);
  const double q[4] = {0, 0, 0, 1}, t[3] = {0, 0, 0};
  double r[3], J[18];
  e.Evaluate(q, t, r, J); f.Evaluate(q, t, r, J); g.Evaluate(q, t, r, J);
  return (r[0] == 0.0 && e.s == 1.0 && f.ljm_norm.x() != 0.0) ? 0 : 1;
}
'''
    subprocess.run(["g++", "-std=c++11", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), "-x", "c++", "-", "-o", "/dev/null"],
                   input=src, text=True, check=True)


@pytest.mark.gpu
def test_blocks_packed_by_the_header_on_the_gpu(cases, aloam, orc):
    """J^T J / J^T r / cost of blocks built through the header: aloam_normal_equations == header Evaluate() + Huber"""
    recs, _ = cases
    ctx = aloam.Aloam(n_scans=16, max_points=4096)
    groups = {}
    for x, b, r, j in recs:
        groups.setdefault(tuple(x), []).append((b, r, j))
    checked = 0
    for x, items in list(groups.items())[:12]:
        x = np.array(x)
        if 0.0 < np.linalg.norm(x[:3]) < 1e-6:
            continue    # ~1e-7 rad with s != 1: the autodiff reference itself is only good to 1e-8 there
        blocks = np.array([b for b, _, _ in items])
        JtJ, Jtr, cost = ctx.normal_equations(blocks, x)
        rJ, rr, rc = orc.normal_equations(blocks, x, huber=0.1, autodiff=True)
        scale = max(1.0, np.abs(rJ).max())
        assert np.abs(JtJ - rJ).max() / scale < 1e-10 and np.abs(Jtr - rr).max() / scale < 1e-10 and abs(cost - rc) < 1e-10 * max(1.0, rc)
        checked += 1
    assert checked >= 6
    ctx.close()
