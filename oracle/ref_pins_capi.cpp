// ref_pins_capi.cpp — builds oracle/_ref/libref_pins.so from the REFERENCE'S OWN statements: the fragments that
// oracle/extract_ref.py cuts, by line range, out of the reference where it lies (sdf_map.cpp:618-715,
// minco.hpp:43-198, optimizer.cpp:573-591 and :1069-1106) are #included below into stand-in classes that declare
// nothing but the members those statements touch.  No reference source is copied into the repository (the fragments
// live under the git-ignored oracle/_ref/gen).  TEST INFRASTRUCTURE ONLY: pins the oracle's restatement of
// updateESDF2d / fillESDF (E1, E2), BandedSystem (M1), RealT2VirtualT / VirtualT2RealT / backwardGradT (M5) and
// positiveSmoothedL1 (P4) to reference-compiled code, bit for bit (tests/test_ref_pins.py).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include <Eigen/Eigen>

#include "../include/alore_b200.h"

using std::ceil;
using std::floor;
using std::sqrt;

// ---- SDFmap: the members updateESDF2d / fillESDF read and write (sdf_map.h:69-183) --------------------------------
class SDFmap {
 public:
  enum { Unknown, Unoccupied, Occupied };   // sdf_map.h:98
  uint8_t* gridmap_ = nullptr;
  std::vector<double> distance_buffer_all_;
  Eigen::Vector2d odom_pos_;
  double detection_range_ = 0.0;
  double global_x_lower_ = 0.0, global_y_lower_ = 0.0, global_x_upper_ = 0.0, global_y_upper_ = 0.0;
  double grid_interval_ = 0.0, inv_grid_interval_ = 0.0;
  int GLX_SIZE_ = 0, GLY_SIZE_ = 0;
  void updateESDF2d();
  template <typename F_get_val, typename F_set_val>
  void fillESDF(F_get_val f_get_val, F_set_val f_set_val, int start, int end, int dim_size);
  void publish_ESDF();
  // the lookup functions of the optimizer's penalty functional and of the collision checks (sdf_map.cpp:453-472, 525-531,
  // 739-871, 942-948)
  Eigen::Vector2d gridIndex2coordd(const Eigen::Vector2i& index);
  Eigen::Vector2d gridIndex2coordd(const int& x, const int& y);
  Eigen::Vector2i coord2gridIndex(const Eigen::Vector2d& pt);
  int Index2Vectornum(const int& x, const int& y);
  int Index2Vectornum(const Eigen::Vector2i& index);
  inline double getDistance(const Eigen::Vector2i& id);
  inline double getDistance(const int& idx, const int& idy);
  inline Eigen::Vector2i ESDFcoord2gridIndex(const Eigen::Vector2d& pt);
  double getDistWithGradBilinear(const Eigen::Vector2d& pos, Eigen::Vector2d& grad);
  double getDistWithGradBilinear(const Eigen::Vector2d& pos, Eigen::Vector2d& grad, const double& mindis);
  double getDistWithGradBilinear(const Eigen::Vector2d& pos);
  double getDistanceReal(const Eigen::Vector2d& pos);
  bool isOccWithSafeDis(const Eigen::Vector2i& index, const double& safe_dis);
  bool isOccWithSafeDis(const int& idx, const int& idy, const double& safe_dis);
};
#include "_ref/gen/ref_sdf_esdf.inc"
#include "_ref/gen/ref_sdf_index.inc"
#include "_ref/gen/ref_sdf_vecnum.inc"
#include "_ref/gen/ref_sdf_lookup.inc"
#include "_ref/gen/ref_sdf_isocc.inc"

// ---- minco::BandedSystem, the whole class --------------------------------------------------------------------------
namespace minco {
#include "_ref/gen/ref_minco_banded.inc"
}

// ---- MSPlanner: the scalar maps ------------------------------------------------------------------------------------
class MSPlanner {
 public:
  double smoothEps = 0.0;
  template <typename EIGENVEC> inline void RealT2VirtualT(const Eigen::VectorXd& RT, EIGENVEC& VT);
  template <typename EIGENVEC> inline void VirtualT2RealT(const EIGENVEC& VT, Eigen::VectorXd& RT);
  inline void positiveSmoothedL1(const double& x, double& f, double& df);
  template <typename EIGENVEC>
  static inline void backwardGradT(const Eigen::VectorXd& tau, const Eigen::VectorXd& gradT, EIGENVEC& gradTau);
};
#include "_ref/gen/ref_opt_tmaps.inc"
#include "_ref/gen/ref_opt_smoothl1.inc"

extern "C" {

// SDFmap::updateESDF2d() exactly as the reference runs it: the window comes from odom_pos_ / detection_range_.
int ref_esdf_update(const alore_map_geom_t* g, const uint8_t* occ, double odom_x, double odom_y, double range, double* dist_inout) {
  SDFmap m;
  m.gridmap_ = const_cast<uint8_t*>(occ);
  m.GLX_SIZE_ = g->glx; m.GLY_SIZE_ = g->gly;
  m.global_x_lower_ = g->x_lower; m.global_y_lower_ = g->y_lower; m.global_x_upper_ = g->x_upper; m.global_y_upper_ = g->y_upper;
  m.grid_interval_ = g->grid_interval; m.inv_grid_interval_ = g->inv_grid_interval;
  m.odom_pos_ = Eigen::Vector2d(odom_x, odom_y);
  m.detection_range_ = range;
  const size_t n = (size_t)g->glx * g->gly;
  m.distance_buffer_all_.assign(dist_inout, dist_inout + n);
  m.updateESDF2d();
  std::memcpy(dist_inout, m.distance_buffer_all_.data(), n * sizeof(double));
  return 0;
}

static void ref_map_init(SDFmap& m, const alore_map_geom_t* g, const double* dist) {
  m.GLX_SIZE_ = g->glx; m.GLY_SIZE_ = g->gly;
  m.global_x_lower_ = g->x_lower; m.global_y_lower_ = g->y_lower; m.global_x_upper_ = g->x_upper; m.global_y_upper_ = g->y_upper;
  m.grid_interval_ = g->grid_interval; m.inv_grid_interval_ = g->inv_grid_interval;
  m.distance_buffer_all_.assign(dist, dist + (size_t)g->glx * g->gly);
}

// n lookups through the reference's own getDistWithGradBilinear (which = 3: with mindis, 2: with gradient, 1: value
// only), getDistanceReal (which = 0) and isOccWithSafeDis at coord2gridIndex(pos) (which = -1; out = 0 / 1).
// grad_io [n][2] is in/out: the reference leaves it untouched on some paths.
int ref_dist_lookups(const alore_map_geom_t* g, const double* dist, int n, const double* pos, int which, double mindis, double* out,
                     double* grad_io) {
  SDFmap m;
  ref_map_init(m, g, dist);
  for (int i = 0; i < n; i++) {
    const Eigen::Vector2d p(pos[2 * i], pos[2 * i + 1]);
    Eigen::Vector2d gr(grad_io ? grad_io[2 * i] : 0.0, grad_io ? grad_io[2 * i + 1] : 0.0);
    if (which == 3) out[i] = m.getDistWithGradBilinear(p, gr, mindis);
    else if (which == 2) out[i] = m.getDistWithGradBilinear(p, gr);
    else if (which == 1) out[i] = m.getDistWithGradBilinear(p);
    else if (which == 0) out[i] = m.getDistanceReal(p);
    else out[i] = m.isOccWithSafeDis(m.coord2gridIndex(p), mindis) ? 1.0 : 0.0;
    if (grad_io) { grad_io[2 * i] = gr.x(); grad_io[2 * i + 1] = gr.y(); }
  }
  return 0;
}

// BandedSystem on a dense row-major N x N matrix (entries outside the band ignored), right-hand side b [N][2].
// mode 0: factorizeLU + solve, 1: factorizeLU + solveAdj.  band_out (may be NULL): (p+q+1)*N factor data.
int ref_banded(int N, int p, int q, const double* A, double* b, int mode, double* band_out) {
  minco::BandedSystem bs;
  bs.create(N, p, q);
  for (int i = 0; i < N; i++)
    for (int j = std::max(0, i - p); j <= std::min(N - 1, i + q); j++) bs(i, j) = A[(size_t)i * N + j];
  bs.factorizeLU();
  Eigen::MatrixXd B = Eigen::MatrixXd::Zero(N, 2);
  for (int i = 0; i < N; i++) { B(i, 0) = b[2 * i]; B(i, 1) = b[2 * i + 1]; }
  if (mode == 0) bs.solve(B);
  else bs.solveAdj(B);
  for (int i = 0; i < N; i++) { b[2 * i] = B(i, 0); b[2 * i + 1] = B(i, 1); }
  if (band_out)
    for (int i = 0; i < N; i++)
      for (int j = std::max(0, i - p); j <= std::min(N - 1, i + q); j++) band_out[(size_t)(i - j + q) * N + j] = bs(i, j);
  bs.destroy();
  return 0;
}

void ref_tmaps(int n, const double* in, int which, const double* gradT, double* out) {
  MSPlanner pl;
  Eigen::VectorXd a(n), o, g(n);
  for (int i = 0; i < n; i++) { a(i) = in[i]; g(i) = gradT ? gradT[i] : 0.0; }
  if (which == 0) pl.RealT2VirtualT(a, o);
  else if (which == 1) pl.VirtualT2RealT(a, o);
  else MSPlanner::backwardGradT(a, g, o);
  for (int i = 0; i < n; i++) out[i] = o(i);
}

void ref_smoothed_l1(double eps, double x, double* f, double* df) {
  MSPlanner pl;
  pl.smoothEps = eps;
  pl.positiveSmoothedL1(x, *f, *df);
}

}  // extern "C"
