// ORACLE (test infrastructure) -- scan-to-map registration of laserMapping.cpp:
//   :142-152  transformAssociateToMap / transformUpdate        :154-163  pointAssociateToMap (double -> float store)
//   :554-729  if map corner > 10 && surf > 50: kd-tree build (:558-559), 2 x { corner 5-NN + line fit (:577-622),
//             surf 5-NN + plane fit (:643-687), ceres::Solve (:712-720) }
// Eigen is not in /root/reference: SelfAdjointEigenSolver<Matrix3d> (:605) and colPivHouseholderQr (:663) are
// replaced by a cyclic Jacobi eigen-solver and a column-pivoted Householder QR with the same mathematical
// result (eigenvalues ascending, unit eigenvectors up to sign; least-squares solution).
#include <algorithm>
#include <cmath>
#include "oracle.h"

namespace orc {

void eig3_sym(const double Ain[9], double evals[3], double V[9]) {
  double A[9];
  for (int i = 0; i < 9; ++i) A[i] = Ain[i];
  for (int i = 0; i < 9; ++i) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
    double dsum = A[0] * A[0] + A[4] * A[4] + A[8] * A[8];
    if (off <= 1e-32 * dsum || off == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double apq = A[3 * p + q];
        if (apq == 0.0) continue;
        double theta = (A[3 * q + q] - A[3 * p + p]) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {  // A <- A * G
          double akp = A[3 * k + p], akq = A[3 * k + q];
          A[3 * k + p] = c * akp - s * akq;
          A[3 * k + q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {  // A <- G^T * A
          double apk = A[3 * p + k], aqk = A[3 * q + k];
          A[3 * p + k] = c * apk - s * aqk;
          A[3 * q + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {  // V <- V * G
          double vkp = V[3 * k + p], vkq = V[3 * k + q];
          V[3 * k + p] = c * vkp - s * vkq;
          V[3 * k + q] = s * vkp + c * vkq;
        }
      }
  }
  double e[3] = {A[0], A[4], A[8]};
  int ord[3] = {0, 1, 2};
  std::sort(ord, ord + 3, [&](int a, int b) { return e[a] < e[b]; });
  double Vs[9];
  for (int j = 0; j < 3; ++j) {
    evals[j] = e[ord[j]];
    for (int k = 0; k < 3; ++k) Vs[3 * k + j] = V[3 * k + ord[j]];
  }
  for (int i = 0; i < 9; ++i) V[i] = Vs[i];
}

void lsq_5x3(const double Ain[15], const double bin[5], double n[3]) {
  const int m = 5, nc = 3;
  double A[15], b[5];
  for (int i = 0; i < 15; ++i) A[i] = Ain[i];
  for (int i = 0; i < 5; ++i) b[i] = bin[i];
  int perm[3] = {0, 1, 2};
  for (int k = 0; k < nc; ++k) {
    // pivot: remaining column with the largest norm
    int best = k; double bestn = -1;
    for (int j = k; j < nc; ++j) {
      double s = 0;
      for (int i = k; i < m; ++i) s += A[3 * i + j] * A[3 * i + j];
      if (s > bestn) { bestn = s; best = j; }
    }
    if (best != k) {
      for (int i = 0; i < m; ++i) std::swap(A[3 * i + k], A[3 * i + best]);
      std::swap(perm[k], perm[best]);
    }
    double nrm = std::sqrt(bestn);
    double akk = A[3 * k + k];
    double alpha = akk > 0 ? -nrm : nrm;
    double v0 = akk - alpha;
    double vn2 = bestn - akk * akk + v0 * v0;
    A[3 * k + k] = v0;
    if (vn2 > 0) {
      for (int j = k + 1; j < nc; ++j) {
        double s = 0;
        for (int i = k; i < m; ++i) s += A[3 * i + k] * A[3 * i + j];
        s = 2.0 * s / vn2;
        for (int i = k; i < m; ++i) A[3 * i + j] -= s * A[3 * i + k];
      }
      double s = 0;
      for (int i = k; i < m; ++i) s += A[3 * i + k] * b[i];
      s = 2.0 * s / vn2;
      for (int i = k; i < m; ++i) b[i] -= s * A[3 * i + k];
    }
    A[3 * k + k] = alpha;
  }
  double y[3];
  for (int k = nc - 1; k >= 0; --k) {
    double s = b[k];
    for (int j = k + 1; j < nc; ++j) s -= A[3 * k + j] * y[j];
    y[k] = s / A[3 * k + k];
  }
  for (int k = 0; k < nc; ++k) n[perm[k]] = y[k];
}

namespace {
// laserMapping.cpp:154-163
inline PointXYZI point_associate_to_map(const PointXYZI& pi, const Quat& q, const Vec3& t) {
  Vec3 w = rotate(q, Vec3{pi.x, pi.y, pi.z}) + t;
  return {(float)w.x, (float)w.y, (float)w.z, pi.intensity};
}
}  // namespace

void Mapping::set_map(const Cloud& c, const Cloud& s) {
  corner_map = c; surf_map = s;
  double t0 = now_ms();
  tree_corner.build(corner_map);
  tree_surf.build(surf_map);
  times.tree_ms = now_ms() - t0;
}

void Mapping::associate(const Cloud& corner_stack, const Cloud& surf_stack, const double x[7],
                        std::vector<MapFit>* fits, std::vector<ResidualBlock>* blocks) const {
  const Quat q{x[0], x[1], x[2], x[3]};
  const Vec3 t{x[4], x[5], x[6]};
  int idx[5]; float sqd[5];
  for (int i = 0; i < (int)corner_stack.size(); ++i) {  // :577-622
    const PointXYZI& ori = corner_stack[i];
    PointXYZI sel = point_associate_to_map(ori, q, t);
    const float qq[3] = {sel.x, sel.y, sel.z};
    if (tree_corner.knn(qq, 5, idx, sqd) < 5) continue;
    if (!(sqd[4] < 1.0)) continue;
    Vec3 near[5], center{0, 0, 0};
    for (int j = 0; j < 5; ++j) {
      near[j] = Vec3{corner_map[idx[j]].x, corner_map[idx[j]].y, corner_map[idx[j]].z};
      center = center + near[j];
    }
    center = Vec3{center.x / 5.0, center.y / 5.0, center.z / 5.0};
    double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < 5; ++j) {
      Vec3 d = near[j] - center;
      const double dv[3] = {d.x, d.y, d.z};
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) cov[3 * a + b] += dv[a] * dv[b];
    }
    double ev[3], V[9];
    eig3_sym(cov, ev, V);
    if (ev[2] > 3 * ev[1]) {  // :611
      Vec3 dir{V[2], V[5], V[8]};
      const double cp[3] = {ori.x, ori.y, ori.z};
      const double a[3] = {0.1 * dir.x + center.x, 0.1 * dir.y + center.y, 0.1 * dir.z + center.z};
      const double b[3] = {-0.1 * dir.x + center.x, -0.1 * dir.y + center.y, -0.1 * dir.z + center.z};
      if (blocks) blocks->push_back(make_edge(cp, a, b, 1.0));
      if (fits) {
        MapFit f; f.query = i; f.type = FACTOR_EDGE; f.d = 0;
        for (int k = 0; k < 3; ++k) { f.p0[k] = a[k]; f.p1[k] = b[k]; }
        for (int k = 0; k < 5; ++k) f.nn[k] = idx[k];
        fits->push_back(f);
      }
    }
  }
  for (int i = 0; i < (int)surf_stack.size(); ++i) {  // :643-687
    const PointXYZI& ori = surf_stack[i];
    PointXYZI sel = point_associate_to_map(ori, q, t);
    const float qq[3] = {sel.x, sel.y, sel.z};
    if (tree_surf.knn(qq, 5, idx, sqd) < 5) continue;
    if (!(sqd[4] < 1.0)) continue;
    double A[15], b[5] = {-1, -1, -1, -1, -1};
    for (int j = 0; j < 5; ++j) {
      A[3 * j] = surf_map[idx[j]].x; A[3 * j + 1] = surf_map[idx[j]].y; A[3 * j + 2] = surf_map[idx[j]].z;
    }
    double nv[3];
    lsq_5x3(A, b, nv);
    const double nn = std::sqrt(nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2]);
    const double negative_OA_dot_norm = 1 / nn;
    nv[0] /= nn; nv[1] /= nn; nv[2] /= nn;
    bool valid = true;
    for (int j = 0; j < 5; ++j) {
      if (std::fabs(nv[0] * surf_map[idx[j]].x + nv[1] * surf_map[idx[j]].y + nv[2] * surf_map[idx[j]].z +
                    negative_OA_dot_norm) > 0.2) { valid = false; break; }
    }
    if (valid) {
      const double cp[3] = {ori.x, ori.y, ori.z};
      if (blocks) blocks->push_back(make_plane_norm(cp, nv, negative_OA_dot_norm));
      if (fits) {
        MapFit f; f.query = i; f.type = FACTOR_PLANE_NORM; f.d = negative_OA_dot_norm;
        for (int k = 0; k < 3; ++k) { f.p0[k] = nv[k]; f.p1[k] = 0; }
        for (int k = 0; k < 5; ++k) f.nn[k] = idx[k];
        fits->push_back(f);
      }
    }
  }
}

int Mapping::register_scan(const Cloud& corner_stack, const Cloud& surf_stack, double x[7], int outer_iters,
                           const SolveOptions& opt) {
  summaries.clear();
  times.assoc_ms = times.solve_ms = 0;
  if (!((int)corner_map.size() > 10 && (int)surf_map.size() > 50)) return 0;  // :554
  for (int it = 0; it < outer_iters; ++it) {
    std::vector<ResidualBlock> blocks;
    double t0 = now_ms();
    associate(corner_stack, surf_stack, x, nullptr, &blocks);
    times.assoc_ms += now_ms() - t0;
    t0 = now_ms();
    SolveSummary S;
    solve(blocks, x, opt, &S);
    summaries.push_back(S);
    times.solve_ms += now_ms() - t0;
  }
  return 1;
}

void transform_associate_to_map(const double qm[4], const double tm[3], const double qo[4], const double to[3],
                                double x[7]) {
  Quat q_wmap_wodom{qm[0], qm[1], qm[2], qm[3]}, q_wodom_curr{qo[0], qo[1], qo[2], qo[3]};
  Quat q = qmul(q_wmap_wodom, q_wodom_curr);
  Vec3 t = rotate(q_wmap_wodom, Vec3{to[0], to[1], to[2]}) + Vec3{tm[0], tm[1], tm[2]};
  x[0] = q.x; x[1] = q.y; x[2] = q.z; x[3] = q.w; x[4] = t.x; x[5] = t.y; x[6] = t.z;
}
void transform_update(const double x[7], const double qo[4], const double to[3], double qm[4], double tm[3]) {
  Quat q_w_curr{x[0], x[1], x[2], x[3]}, q_wodom_curr{qo[0], qo[1], qo[2], qo[3]};
  Quat q = qmul(q_w_curr, qinv(q_wodom_curr));
  Vec3 r = rotate(q, Vec3{to[0], to[1], to[2]});
  qm[0] = q.x; qm[1] = q.y; qm[2] = q.z; qm[3] = q.w;
  tm[0] = x[4] - r.x; tm[1] = x[5] - r.y; tm[2] = x[6] - r.z;
}

}  // namespace orc
