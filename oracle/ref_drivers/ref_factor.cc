// TEST INFRASTRUCTURE.  Compiles the reference's own lidarFactor.hpp (from /root/reference, never copied) against the stand-in
// headers of oracle/ref_shim and exposes its three cost functions through the reference's own factories (Create):
// residuals, and Jacobians with respect to the ambient parameters (q = x, y, z, w ; t) by forward-mode autodiff over the
// reference's templated operator().  Built into oracle/_ref/ by `make -C oracle ref`.
#include REF_LIDAR_FACTOR_HPP

extern "C" {

// type 0 LidarEdgeFactor       pts = curr, last_a, last_b          extra = s
//      1 LidarPlaneFactor      pts = curr, last_j, last_l, last_m  extra = s
//      2 LidarPlaneNormFactor  pts = curr, plane_unit_norm         extra = negative_OA_dot_norm
// residuals[3] (1 used for types 1, 2), jq[rows x 4], jt[rows x 3] row-major (either may be NULL).  Returns the number of rows.
int ref_factor_eval(int type, const double* pts, double extra, const double* q, const double* t, double* residuals, double* jq, double* jt) {
  auto V = [&](int i) { return Eigen::Vector3d(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]); };
  ceres::CostFunction* f = nullptr;
  if (type == 0) f = LidarEdgeFactor::Create(V(0), V(1), V(2), extra);
  else if (type == 1) f = LidarPlaneFactor::Create(V(0), V(1), V(2), V(3), extra);
  else if (type == 2) f = LidarPlaneNormFactor::Create(V(0), V(1), extra);
  else return -1;
  const double* params[2] = {q, t};
  double* jac[2] = {jq, jt};
  const bool ok = f->Evaluate(params, residuals, (jq || jt) ? jac : nullptr);
  const int rows = f->num_residuals();
  delete f;
  return ok ? rows : -2;
}

}  // extern "C"
