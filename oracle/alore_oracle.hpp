// alore_oracle.hpp — CPU ORACLE (test infrastructure, NOT the product).
//
// A literal, dependency-free C++17 restatement of the reference's CPU algorithm for the
// planning_ddr_opt hot path.  It exists only to check the CUDA path (tests/, smoke(),
// bench.py's cpu_baseline / --impl reference leg).  Nothing in the product may include it.
//
// PARITY UNPINNED (floating-point part): the reference ships no tests / golden vectors for
// this path and cannot be compiled here (needs ROS + Eigen3 + PCL; none installed, no
// network).  Eigen 3.3.x supplies the *summation order* of VectorXd::dot/norm/sum and of the
// small mat-vec products; this restatement fixes it to sequential-in-index.  The ESDF part
// uses no Eigen arithmetic, so its bit-exact parity is well defined and is pinned against an
// independent brute-force EDT in tests/.  The L-BFGS part is additionally checked against
// the reference's own lbfgs.hpp compiled through a tiny Eigen stand-in (oracle/_ref).
//
// Citations: paths relative to /root/reference/planning_ddr_opt/
//   sdf   = utils/plan_env/src/sdf_map.cpp           opt  = back_end/src/optimizer.cpp
//   minco = back_end/include/gcopter/minco.hpp       lbf  = back_end/include/gcopter/lbfgs.hpp
//   traj  = back_end/include/gcopter/trajectory.hpp  jps  = front_end/src/jps_planner/jps_planner.cpp
#pragma once
#include <algorithm>
#include <array>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "../include/alore_b200.h"  // POD parameter / geometry structs only

namespace orc {

// TEST HOOK, default off.  The reference pushes a collision sample's position gradient onto the chain of
// every sample up to AND INCLUDING itself and then weights it with the full-piece Simpson weights
// (optimizer.cpp:945-946, 1056-1057); the sample's own weight in the partial integral up to its time is
// 0 (j = 0), 1 (interior even j, where the full-piece weight is 2) or 1 (j = 2K), so the reference's collision
// gradient is not the exact derivative of its own cost.  With this flag the oracle uses the exact own-weight,
// which lets tests check EVERY other term of the restatement against finite differences.
inline bool g_exact_chain_weights = false;

// sin/cos used by the penalty functionals and check_final_collision.  Default: glibc (what the reference
// calls).  `g_trig_portable`: a self-contained fdlibm-style evaluation (3-term Cody-Waite reduction by pi/2 +
// the classic degree-13/14 kernels) built only from IEEE +,-,*,/ and floor, so that the CUDA kernels — which
// carry their own, independently written copy of the same published algorithm — produce the SAME BITS.
// The reference's optimizer amplifies a 1e-15 input perturbation to percent-level changes of the optimised
// trajectory (measured), so trajectory-level parity is only meaningful under identical
// arithmetic; libm-vs-portable differences are <= 1 ulp per call and are reported by the tests.
inline bool g_trig_portable = false;

namespace ptrig {
constexpr double invpio2 = 6.36619772367581382433e-01;
constexpr double pio2_1 = 1.57079632673412561417e+00, pio2_1t = 6.07710050650619224932e-11;
constexpr double pio2_2 = 6.07710050630396597660e-11, pio2_2t = 2.02226624879595063154e-21;
constexpr double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
                 S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
constexpr double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
                 C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
inline double ksin(double x, double y) {
  const double z = x * x;
  const double v = z * x;
  const double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
  return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}
inline double kcos(double x, double y) {
  const double z = x * x;
  const double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
  const double ax = x < 0.0 ? -x : x;
  if (ax < 0.3) return 1.0 - (0.5 * z - (z * r - x * y));
  const double qx = ax > 0.78125 ? 0.28125 : std::floor(0.25 * ax * 4194304.0) / 4194304.0;
  const double hz = 0.5 * z - qx;
  const double a = 1.0 - qx;
  return a - (hz - (z * r - x * y));
}
inline void sincos(double x, double& s, double& c) {
  if (!(x > -1.0e5 && x < 1.0e5)) { s = std::sin(x); c = std::cos(x); return; }
  const double fn = std::floor(x * invpio2 + 0.5);
  const int n = (int)fn;
  // fdlibm __ieee754_rem_pio2, medium-size path, second iteration taken unconditionally (118-bit pi/2):
  // with |x| < 1e5 the closest any double gets to a multiple of pi/2 still leaves > 60 significant bits.
  double r = x - fn * pio2_1;
  const double t = r;
  double w = fn * pio2_2;
  r = t - w;
  w = fn * pio2_2t - ((t - r) - w);
  const double y0 = r - w;
  const double y1 = (r - y0) - w;
  const double ks = ksin(y0, y1), kc = kcos(y0, y1);
  switch (n & 3) {
    case 0: s = ks; c = kc; break;
    case 1: s = kc; c = -ks; break;
    case 2: s = -ks; c = -kc; break;
    default: s = -kc; c = ks; break;
  }
}
}  // namespace ptrig
inline double osin(double x) { if (!g_trig_portable) return std::sin(x); double s, c; ptrig::sincos(x, s, c); return s; }
inline double ocos(double x) { if (!g_trig_portable) return std::cos(x); double s, c; ptrig::sincos(x, s, c); return c; }

// =====================================================================================
// E1-E4  grid map + ESDF                                     sdf:453-472, 618-715, 739-871
// =====================================================================================
struct SdfMap {
  alore_map_geom_t g{};
  std::vector<uint8_t> gridmap_;              // sdf_map.h:73
  std::vector<double> distance_buffer_all_;   // sdf_map.h:69, ctor fills DBL_MAX (sdf_map.h:160)
  const double* dist_view_ = nullptr;         // capi: read-only view on a caller-owned distance buffer
  const double* dist() const { return dist_view_ ? dist_view_ : distance_buffer_all_.data(); }

  void init(const alore_map_geom_t& geom) {
    g = geom;
    gridmap_.assign(size_t(g.glx) * g.gly, ALORE_UNKNOWN);
    distance_buffer_all_.assign(size_t(g.glx) * g.gly, std::numeric_limits<double>::max());
  }

  // sdf:618-621 — window corners from odom and detection range (FP then truncation, quirk q4).
  void window(double odom_x, double odom_y, double detection_range, int mn[2], int mx[2]) const {
    mn[0] = int(std::floor(std::max(0.0, odom_x - detection_range - g.x_lower) * g.inv_grid_interval));
    mn[1] = int(std::floor(std::max(0.0, odom_y - detection_range - g.y_lower) * g.inv_grid_interval));
    mx[0] = int(std::ceil(std::min(g.x_upper - g.x_lower, odom_x + detection_range - g.x_lower) * g.inv_grid_interval) - 1);
    mx[1] = int(std::ceil(std::min(g.y_upper - g.y_lower, odom_y + detection_range - g.y_lower) * g.inv_grid_interval) - 1);
  }

  // sdf:682-715 — 1-D squared-distance lower envelope, DBL_MAX as infinity, int q*q.
  template <typename FG, typename FS>
  static void fillESDF(FG f_get_val, FS f_set_val, int start, int end, int dim_size) {
    std::vector<int> v(dim_size);
    std::vector<double> z(dim_size + 1);
    int k = start;
    v[start] = start;
    z[start] = -std::numeric_limits<double>::max();
    z[start + 1] = std::numeric_limits<double>::max();
    for (int q = start + 1; q <= end; q++) {
      k++;
      double s;
      do {
        k--;
        s = ((f_get_val(q) + q * q) - (f_get_val(v[k]) + v[k] * v[k])) / (2 * q - 2 * v[k]);
      } while (s <= z[k]);
      k++;
      v[k] = q;
      z[k] = s;
      z[k + 1] = std::numeric_limits<double>::max();
    }
    k = start;
    for (int q = start; q <= end; q++) {
      while (z[k + 1] < q) k++;
      double val = (q - v[k]) * (q - v[k]) + f_get_val(v[k]);
      f_set_val(q, val);
    }
  }

  // sdf:618-680 — literal, including the x*update_Y_SIZE aliasing stride and the `<` combine
  // bounds.  If sq_pos/sq_neg are given they receive the pre-sqrt values `val` exactly as the
  // reference's second pass sees them, at the reference's own buffer index x*update_Y_SIZE+y
  // (size (X+1)*(Y+1)), so the GPU's integer squared distances can be compared bit for bit.
  void updateESDF2d(const int min_esdf[2], const int max_esdf[2],
                    std::vector<double>* sq_pos = nullptr, std::vector<double>* sq_neg = nullptr) {
    updateESDF2d(gridmap_.data(), distance_buffer_all_.data(), min_esdf, max_esdf, sq_pos, sq_neg);
  }
  void updateESDF2d(const uint8_t* grid, double* dall, const int min_esdf[2], const int max_esdf[2],
                    std::vector<double>* sq_pos, std::vector<double>* sq_neg) const {
    const int GLY = g.gly;
    const double grid_interval_ = g.grid_interval;
    int update_X_SIZE = max_esdf[0] - min_esdf[0];
    int update_Y_SIZE = max_esdf[1] - min_esdf[1];
    int update_XY_SIZE = (update_X_SIZE + 1) * (update_Y_SIZE + 1);
    std::vector<double> tmp_buffer1_(update_XY_SIZE, 0.0);
    std::vector<double> distance_buffer_(update_XY_SIZE, 0.0);
    std::vector<double> distance_buffer_neg_(update_XY_SIZE, 0.0);
    if (sq_pos) sq_pos->assign(update_XY_SIZE, 0.0);
    if (sq_neg) sq_neg->assign(update_XY_SIZE, 0.0);
    const double DMAX = std::numeric_limits<double>::max();
    for (int x = 0; x <= update_X_SIZE; x++) {
      fillESDF(
          [&](int y) {
            return grid[size_t(x + min_esdf[0]) * GLY + (y + min_esdf[1])] == ALORE_OCCUPIED ? 0.0 : DMAX;
          },
          [&](int y, double val) { tmp_buffer1_[x * update_Y_SIZE + y] = val; }, 0, update_Y_SIZE,
          update_Y_SIZE + 1);
    }
    for (int y = 0; y <= update_Y_SIZE; y++) {
      fillESDF([&](int x) { return tmp_buffer1_[x * update_Y_SIZE + y]; },
               [&](int x, double val) {
                 distance_buffer_[x * update_Y_SIZE + y] = grid_interval_ * std::sqrt(val);
                 if (sq_pos) (*sq_pos)[x * update_Y_SIZE + y] = val;
               },
               0, update_X_SIZE, update_X_SIZE + 1);
    }
    for (int x = 0; x <= update_X_SIZE; x++) {
      fillESDF(
          [&](int y) {
            int state = grid[size_t(x + min_esdf[0]) * GLY + (y + min_esdf[1])];
            return (state == ALORE_UNOCCUPIED || state == ALORE_UNKNOWN) ? 0.0 : DMAX;
          },
          [&](int y, double val) { tmp_buffer1_[x * update_Y_SIZE + y] = val; }, 0, update_Y_SIZE,
          update_Y_SIZE + 1);
    }
    for (int y = 0; y <= update_Y_SIZE; y++) {
      fillESDF([&](int x) { return tmp_buffer1_[x * update_Y_SIZE + y]; },
               [&](int x, double val) {
                 distance_buffer_neg_[x * update_Y_SIZE + y] = grid_interval_ * std::sqrt(val);
                 if (sq_neg) (*sq_neg)[x * update_Y_SIZE + y] = val;
               },
               0, update_X_SIZE, update_X_SIZE + 1);
    }
    for (int x = 0; x < update_X_SIZE; x++)
      for (int y = 0; y < update_Y_SIZE; y++) {
        size_t global_idx = size_t(x + min_esdf[0]) * GLY + y + min_esdf[1];
        int idx = x * update_Y_SIZE + y;
        dall[global_idx] = distance_buffer_[idx];
        if (distance_buffer_neg_[idx] > 0.0)
          dall[global_idx] += (-distance_buffer_neg_[idx] + grid_interval_);
      }
  }

  // sdf:753-758
  void ESDFcoord2gridIndex(const double pt[2], int idx[2]) const {
    idx[0] = std::min(std::max(int((pt[0] - g.x_lower) * g.inv_grid_interval - 0.5), 0), g.glx - 1);
    idx[1] = std::min(std::max(int((pt[1] - g.y_lower) * g.inv_grid_interval - 0.5), 0), g.gly - 1);
  }
  // sdf:453-465
  void gridIndex2coordd(const int idx[2], double pt[2]) const {
    pt[0] = ((double)idx[0] + 0.5) * g.grid_interval + g.x_lower;
    pt[1] = ((double)idx[1] + 0.5) * g.grid_interval + g.y_lower;
  }
  // sdf:467-472
  void coord2gridIndex(const double pt[2], int idx[2]) const {
    idx[0] = std::min(std::max(int((pt[0] - g.x_lower) * g.inv_grid_interval), 0), g.glx - 1);
    idx[1] = std::min(std::max(int((pt[1] - g.y_lower) * g.inv_grid_interval), 0), g.gly - 1);
  }
  double getDistance(int ix, int iy) const { return dist()[size_t(ix) * g.gly + iy]; }

  bool outOfMap(const double pos[2]) const {
    return pos[0] < g.x_lower || pos[1] < g.y_lower || pos[0] > g.x_upper || pos[1] > g.y_upper;
  }

  // sdf:796-834 — 3-argument overload: 1e10 outside, early return WITHOUT touching grad when
  // dist > mindis.
  double getDistWithGradBilinear(const double pos[2], double grad[2], double mindis) const {
    if (outOfMap(pos)) { grad[0] = 0.0; grad[1] = 0.0; return 1e10; }
    int idx[2];
    ESDFcoord2gridIndex(pos, idx);
    if (idx[0] >= g.glx - 1 || idx[1] >= g.gly - 1) { grad[0] = 0.0; grad[1] = 0.0; return 1e10; }
    double idx_pos[2];
    gridIndex2coordd(idx, idx_pos);
    double diff[2] = {(pos[0] - idx_pos[0]) * g.inv_grid_interval, (pos[1] - idx_pos[1]) * g.inv_grid_interval};
    double values[2][2];
    for (int x = 0; x < 2; x++)
      for (int y = 0; y < 2; y++) values[x][y] = getDistance(idx[0] + x, idx[1] + y);
    double v0 = (1 - diff[0]) * values[0][0] + diff[0] * values[1][0];
    double v1 = (1 - diff[0]) * values[0][1] + diff[0] * values[1][1];
    double dist = (1 - diff[1]) * v0 + diff[1] * v1;
    if (dist > mindis) return dist;
    grad[1] = (v1 - v0) * g.inv_grid_interval;
    grad[0] = ((1 - diff[1]) * (values[1][0] - values[0][0]) + diff[1] * (values[1][1] - values[0][1])) * g.inv_grid_interval;
    return dist;
  }
  // sdf:760-794 — 2-argument overload: 100 outside, grad always written.
  double getDistWithGradBilinear(const double pos[2], double grad[2]) const {
    if (outOfMap(pos)) { grad[0] = 0.0; grad[1] = 0.0; return 100; }
    int idx[2];
    ESDFcoord2gridIndex(pos, idx);
    if (idx[0] >= g.glx - 1 || idx[1] >= g.gly - 1) { grad[0] = 0.0; grad[1] = 0.0; return 100; }
    double idx_pos[2];
    gridIndex2coordd(idx, idx_pos);
    double diff[2] = {(pos[0] - idx_pos[0]) * g.inv_grid_interval, (pos[1] - idx_pos[1]) * g.inv_grid_interval};
    double values[2][2];
    for (int x = 0; x < 2; x++)
      for (int y = 0; y < 2; y++) values[x][y] = getDistance(idx[0] + x, idx[1] + y);
    double v0 = (1 - diff[0]) * values[0][0] + diff[0] * values[1][0];
    double v1 = (1 - diff[0]) * values[0][1] + diff[0] * values[1][1];
    double dist = (1 - diff[1]) * v0 + diff[1] * v1;
    grad[1] = (v1 - v0) * g.inv_grid_interval;
    grad[0] = ((1 - diff[1]) * (values[1][0] - values[0][0]) + diff[1] * (values[1][1] - values[0][1])) * g.inv_grid_interval;
    return dist;
  }
  // sdf:836-863 — 1-argument overload.
  double getDistWithGradBilinear(const double pos[2]) const {
    if (outOfMap(pos)) return 1e10;
    int idx[2];
    ESDFcoord2gridIndex(pos, idx);
    if (idx[0] >= g.glx - 1 || idx[1] >= g.gly - 1) return 1e10;
    double idx_pos[2];
    gridIndex2coordd(idx, idx_pos);
    double diff[2] = {(pos[0] - idx_pos[0]) * g.inv_grid_interval, (pos[1] - idx_pos[1]) * g.inv_grid_interval};
    double values[2][2];
    for (int x = 0; x < 2; x++)
      for (int y = 0; y < 2; y++) values[x][y] = getDistance(idx[0] + x, idx[1] + y);
    double v0 = (1 - diff[0]) * values[0][0] + diff[0] * values[1][0];
    double v1 = (1 - diff[0]) * values[0][1] + diff[0] * values[1][1];
    return (1 - diff[1]) * v0 + diff[1] * v1;
  }
  // sdf:865-871
  double getDistanceReal(const double pos[2]) const {
    if (outOfMap(pos)) return 10000;
    int idx[2];
    coord2gridIndex(pos, idx);
    return dist()[size_t(idx[0]) * g.gly + idx[1]];
  }
  // sdf:942-948
  bool isOccWithSafeDis(int ix, int iy, double safe_dis) const {
    return dist()[size_t(ix) * g.gly + iy] < safe_dis;
  }
};

// =====================================================================================
// M1  BandedSystem                                                          minco:43-198
// =====================================================================================
struct BandedSystem {
  int N = 0, lowerBw = 0, upperBw = 0;
  std::vector<double> ptrData;
  void create(int n, int p, int q) {
    N = n; lowerBw = p; upperBw = q;
    ptrData.assign(size_t(N) * (lowerBw + upperBw + 1), 0.0);
  }
  void reset() { std::fill(ptrData.begin(), ptrData.end(), 0.0); }
  double& operator()(int i, int j) { return ptrData[size_t(i - j + upperBw) * N + j]; }
  const double& operator()(int i, int j) const { return ptrData[size_t(i - j + upperBw) * N + j]; }

  void factorizeLU() {  // minco:99-131, no pivoting, zero-skips kept
    int iM, jM;
    double cVl;
    for (int k = 0; k <= N - 2; k++) {
      iM = std::min(k + lowerBw, N - 1);
      cVl = operator()(k, k);
      for (int i = k + 1; i <= iM; i++)
        if (operator()(i, k) != 0.0) operator()(i, k) /= cVl;
      jM = std::min(k + upperBw, N - 1);
      for (int j = k + 1; j <= jM; j++) {
        cVl = operator()(k, j);
        if (cVl != 0.0)
          for (int i = k + 1; i <= iM; i++)
            if (operator()(i, k) != 0.0) operator()(i, j) -= operator()(i, k) * cVl;
      }
    }
  }
  // b is N x 2 row-major (b[2*i+c]).                                        minco:137-164
  void solve(double* b) const {
    int iM;
    for (int j = 0; j <= N - 1; j++) {
      iM = std::min(j + lowerBw, N - 1);
      for (int i = j + 1; i <= iM; i++)
        if (operator()(i, j) != 0.0) {
          double a = operator()(i, j);
          b[2 * i] -= a * b[2 * j];
          b[2 * i + 1] -= a * b[2 * j + 1];
        }
    }
    for (int j = N - 1; j >= 0; j--) {
      double d = operator()(j, j);
      b[2 * j] /= d;
      b[2 * j + 1] /= d;
      iM = std::max(0, j - upperBw);
      for (int i = iM; i <= j - 1; i++)
        if (operator()(i, j) != 0.0) {
          double a = operator()(i, j);
          b[2 * i] -= a * b[2 * j];
          b[2 * i + 1] -= a * b[2 * j + 1];
        }
    }
  }
  void solveAdj(double* b) const {  // minco:170-197
    int iM;
    for (int j = 0; j <= N - 1; j++) {
      double d = operator()(j, j);
      b[2 * j] /= d;
      b[2 * j + 1] /= d;
      iM = std::min(j + upperBw, N - 1);
      for (int i = j + 1; i <= iM; i++)
        if (operator()(j, i) != 0.0) {
          double a = operator()(j, i);
          b[2 * i] -= a * b[2 * j];
          b[2 * i + 1] -= a * b[2 * j + 1];
        }
    }
    for (int j = N - 1; j >= 0; j--) {
      iM = std::max(0, j - lowerBw);
      for (int i = iM; i <= j - 1; i++)
        if (operator()(j, i) != 0.0) {
          double a = operator()(j, i);
          b[2 * i] -= a * b[2 * j];
          b[2 * i + 1] -= a * b[2 * j + 1];
        }
    }
  }
};

// =====================================================================================
// M2-M4  MINCO_S3NU                                                      minco:751-1209
//   headPVA/tailPVA are 2x3: [dim][P,V,A]; b is 6N x 2 row-major, column 0 = yaw, 1 = s.
// =====================================================================================
struct MincoS3NU {
  int N = 0;
  double headPVA[2][3]{}, tailPVA[2][3]{};
  BandedSystem A;
  std::vector<double> b, T1, T2, T3, T4, T5;
  double energyWeights[2] = {1.0, 1.0};

  void setConditions(const double head[2][3], const double tail[2][3], int pieceNum, const double ew[2]) {
    N = pieceNum;
    std::memcpy(headPVA, head, sizeof(headPVA));
    std::memcpy(tailPVA, tail, sizeof(tailPVA));
    A.create(6 * N, 6, 6);
    b.assign(size_t(12) * N, 0.0);
    T1.assign(N, 0.0); T2 = T1; T3 = T1; T4 = T1; T5 = T1;
    energyWeights[0] = ew[0]; energyWeights[1] = ew[1];
  }
  void setTConditions(const double tail[2][3]) { std::memcpy(tailPVA, tail, sizeof(tailPVA)); }
  double& B(int r, int c) { return b[2 * size_t(r) + c]; }
  double B(int r, int c) const { return b[2 * size_t(r) + c]; }

  // inPs: 2 x (N-1) column-major like Eigen (inPs[2*i+dim]); ts: N.           minco:817-898
  void setParameters(const double* inPs, const double* ts) {
    for (int i = 0; i < N; i++) {
      T1[i] = ts[i];
      T2[i] = T1[i] * T1[i];
      T3[i] = T2[i] * T1[i];
      T4[i] = T2[i] * T2[i];
      T5[i] = T4[i] * T1[i];
    }
    A.reset();
    std::fill(b.begin(), b.end(), 0.0);
    A(0, 0) = 1.0; A(1, 1) = 1.0; A(2, 2) = 2.0;
    for (int d = 0; d < 2; d++) { B(0, d) = headPVA[d][0]; B(1, d) = headPVA[d][1]; B(2, d) = headPVA[d][2]; }
    for (int i = 0; i < N - 1; i++) {
      A(6 * i + 3, 6 * i + 3) = 6.0;
      A(6 * i + 3, 6 * i + 4) = 24.0 * T1[i];
      A(6 * i + 3, 6 * i + 5) = 60.0 * T2[i];
      A(6 * i + 3, 6 * i + 9) = -6.0;
      A(6 * i + 4, 6 * i + 4) = 24.0;
      A(6 * i + 4, 6 * i + 5) = 120.0 * T1[i];
      A(6 * i + 4, 6 * i + 10) = -24.0;
      A(6 * i + 5, 6 * i) = 1.0;
      A(6 * i + 5, 6 * i + 1) = T1[i];
      A(6 * i + 5, 6 * i + 2) = T2[i];
      A(6 * i + 5, 6 * i + 3) = T3[i];
      A(6 * i + 5, 6 * i + 4) = T4[i];
      A(6 * i + 5, 6 * i + 5) = T5[i];
      A(6 * i + 6, 6 * i) = 1.0;
      A(6 * i + 6, 6 * i + 1) = T1[i];
      A(6 * i + 6, 6 * i + 2) = T2[i];
      A(6 * i + 6, 6 * i + 3) = T3[i];
      A(6 * i + 6, 6 * i + 4) = T4[i];
      A(6 * i + 6, 6 * i + 5) = T5[i];
      A(6 * i + 6, 6 * i + 6) = -1.0;
      A(6 * i + 7, 6 * i + 1) = 1.0;
      A(6 * i + 7, 6 * i + 2) = 2 * T1[i];
      A(6 * i + 7, 6 * i + 3) = 3 * T2[i];
      A(6 * i + 7, 6 * i + 4) = 4 * T3[i];
      A(6 * i + 7, 6 * i + 5) = 5 * T4[i];
      A(6 * i + 7, 6 * i + 7) = -1.0;
      A(6 * i + 8, 6 * i + 2) = 2.0;
      A(6 * i + 8, 6 * i + 3) = 6 * T1[i];
      A(6 * i + 8, 6 * i + 4) = 12 * T2[i];
      A(6 * i + 8, 6 * i + 5) = 20 * T3[i];
      A(6 * i + 8, 6 * i + 8) = -2.0;
      B(6 * i + 5, 0) = inPs[2 * i];
      B(6 * i + 5, 1) = inPs[2 * i + 1];
    }
    A(6 * N - 3, 6 * N - 6) = 1.0;
    A(6 * N - 3, 6 * N - 5) = T1[N - 1];
    A(6 * N - 3, 6 * N - 4) = T2[N - 1];
    A(6 * N - 3, 6 * N - 3) = T3[N - 1];
    A(6 * N - 3, 6 * N - 2) = T4[N - 1];
    A(6 * N - 3, 6 * N - 1) = T5[N - 1];
    A(6 * N - 2, 6 * N - 5) = 1.0;
    A(6 * N - 2, 6 * N - 4) = 2 * T1[N - 1];
    A(6 * N - 2, 6 * N - 3) = 3 * T2[N - 1];
    A(6 * N - 2, 6 * N - 2) = 4 * T3[N - 1];
    A(6 * N - 2, 6 * N - 1) = 5 * T4[N - 1];
    A(6 * N - 1, 6 * N - 4) = 2;
    A(6 * N - 1, 6 * N - 3) = 6 * T1[N - 1];
    A(6 * N - 1, 6 * N - 2) = 12 * T2[N - 1];
    A(6 * N - 1, 6 * N - 1) = 20 * T3[N - 1];
    for (int d = 0; d < 2; d++) {
      B(6 * N - 3, d) = tailPVA[d][0];
      B(6 * N - 2, d) = tailPVA[d][1];
      B(6 * N - 1, d) = tailPVA[d][2];
    }
    A.factorizeLU();
    A.solve(b.data());
  }

  // (row * W).dot(row') with W = diag(energyWeights): sequential over the 2 dims.
  double wdot(int r1, int r2) const {
    return (B(r1, 0) * energyWeights[0]) * B(r2, 0) + (B(r1, 1) * energyWeights[1]) * B(r2, 1);
  }
  void getEnergy(double& energy) const {  // minco:915-934
    energy = 0.0;
    for (int i = 0; i < N; i++) {
      energy += 36.0 * wdot(6 * i + 3, 6 * i + 3) * T1[i] +
                144.0 * wdot(6 * i + 4, 6 * i + 3) * T2[i] +
                192.0 * wdot(6 * i + 4, 6 * i + 4) * T3[i] +
                240.0 * wdot(6 * i + 5, 6 * i + 3) * T3[i] +
                720.0 * wdot(6 * i + 5, 6 * i + 4) * T4[i] +
                720.0 * wdot(6 * i + 5, 6 * i + 5) * T5[i];
    }
  }
  void getEnergyPartialGradByCoeffs(std::vector<double>& gdC) const {  // minco:941-969
    gdC.assign(size_t(12) * N, 0.0);
    for (int i = 0; i < N; i++)
      for (int d = 0; d < 2; d++) {
        const double w = energyWeights[d];
        gdC[2 * (6 * i + 5) + d] = 240.0 * B(6 * i + 3, d) * w * T3[i] + 720.0 * B(6 * i + 4, d) * w * T4[i] + 1440.0 * B(6 * i + 5, d) * w * T5[i];
        gdC[2 * (6 * i + 4) + d] = 144.0 * B(6 * i + 3, d) * w * T2[i] + 384.0 * B(6 * i + 4, d) * w * T3[i] + 720.0 * B(6 * i + 5, d) * w * T4[i];
        gdC[2 * (6 * i + 3) + d] = 72.0 * B(6 * i + 3, d) * w * T1[i] + 144.0 * B(6 * i + 4, d) * w * T2[i] + 240.0 * B(6 * i + 5, d) * w * T3[i];
      }
  }
  void getEnergyPartialGradByTimes(std::vector<double>& gdT) const {  // minco:971-992
    gdT.assign(N, 0.0);
    for (int i = 0; i < N; i++) {
      gdT[i] = 36.0 * wdot(6 * i + 3, 6 * i + 3) +
               288.0 * wdot(6 * i + 4, 6 * i + 3) * T1[i] +
               576.0 * wdot(6 * i + 4, 6 * i + 4) * T2[i] +
               720.0 * wdot(6 * i + 5, 6 * i + 3) * T2[i] +
               2880.0 * wdot(6 * i + 5, 6 * i + 4) * T3[i] +
               3600.0 * wdot(6 * i + 5, 6 * i + 5) * T4[i];
    }
  }
  // minco:1139-1209.  gradByPoints: 2 x (N-1) column-major; gradByTailStateS: 2.
  void propogateArcYawLenghGrad(const std::vector<double>& partialGradByCoeffs,
                                const std::vector<double>& partialGradByTimes,
                                std::vector<double>& gradByPoints, std::vector<double>& gradByTimes,
                                double gradByTailStateS[2]) const {
    gradByPoints.assign(size_t(2) * std::max(N - 1, 0), 0.0);
    gradByTimes.assign(N, 0.0);
    std::vector<double> adjGrad = partialGradByCoeffs;
    A.solveAdj(adjGrad.data());
    for (int i = 0; i < N - 1; i++) {
      gradByPoints[2 * i] = adjGrad[2 * (6 * i + 5)];
      gradByPoints[2 * i + 1] = adjGrad[2 * (6 * i + 5) + 1];
    }
    double B1[6][2], B2[3][2];
    for (int i = 0; i < N - 1; i++) {
      for (int d = 0; d < 2; d++) {
        B1[2][d] = -(B(i * 6 + 1, d) + 2.0 * T1[i] * B(i * 6 + 2, d) + 3.0 * T2[i] * B(i * 6 + 3, d) +
                     4.0 * T3[i] * B(i * 6 + 4, d) + 5.0 * T4[i] * B(i * 6 + 5, d));
        B1[3][d] = B1[2][d];
        B1[4][d] = -(2.0 * B(i * 6 + 2, d) + 6.0 * T1[i] * B(i * 6 + 3, d) + 12.0 * T2[i] * B(i * 6 + 4, d) +
                     20.0 * T3[i] * B(i * 6 + 5, d));
        B1[5][d] = -(6.0 * B(i * 6 + 3, d) + 24.0 * T1[i] * B(i * 6 + 4, d) + 60.0 * T2[i] * B(i * 6 + 5, d));
        B1[0][d] = -(24.0 * B(i * 6 + 4, d) + 120.0 * T1[i] * B(i * 6 + 5, d));
        B1[1][d] = -120.0 * B(i * 6 + 5, d);
      }
      // B1.cwiseProduct(adj.block<6,2>(6i+3,0)).sum(): column-major traversal of the 6x2 block.
      double s = 0.0;
      for (int d = 0; d < 2; d++)
        for (int r = 0; r < 6; r++) s += B1[r][d] * adjGrad[2 * (6 * i + 3 + r) + d];
      gradByTimes[i] = s;
    }
    for (int d = 0; d < 2; d++) {
      B2[0][d] = -(B(6 * N - 5, d) + 2.0 * T1[N - 1] * B(6 * N - 4, d) + 3.0 * T2[N - 1] * B(6 * N - 3, d) +
                   4.0 * T3[N - 1] * B(6 * N - 2, d) + 5.0 * T4[N - 1] * B(6 * N - 1, d));
      B2[1][d] = -(2.0 * B(6 * N - 4, d) + 6.0 * T1[N - 1] * B(6 * N - 3, d) + 12.0 * T2[N - 1] * B(6 * N - 2, d) +
                   20.0 * T3[N - 1] * B(6 * N - 1, d));
      B2[2][d] = -(6.0 * B(6 * N - 3, d) + 24.0 * T1[N - 1] * B(6 * N - 2, d) + 60.0 * T2[N - 1] * B(6 * N - 1, d));
    }
    {
      double s = 0.0;
      for (int d = 0; d < 2; d++)
        for (int r = 0; r < 3; r++) s += B2[r][d] * adjGrad[2 * (6 * N - 3 + r) + d];
      gradByTimes[N - 1] = s;
    }
    for (int i = 0; i < N; i++) gradByTimes[i] += partialGradByTimes[i];
    gradByTailStateS[0] = adjGrad[2 * (6 * N - 3)];
    gradByTailStateS[1] = adjGrad[2 * (6 * N - 3) + 1];
  }
};

// =====================================================================================
// trajectory.hpp subset used by check_final_collision            traj:75-103, 472-503
// Piece i: coefficient of t^k for dim d is coef[(6*i+k)*2+d] (ascending, as MINCO's b).
// Piece::getPos iterates from the constant term upwards with a running power.
// =====================================================================================
struct Traj {
  int N = 0;
  std::vector<double> T, coef;
  int locatePieceIdx(double& t) const {  // traj:472-490
    int idx;
    double dur;
    for (idx = 0; idx < N && t > (dur = T[idx]); idx++) t -= dur;
    if (idx == N) { idx--; t += T[idx]; }
    return idx;
  }
  void getPos(double t, double pos[2]) const {
    int i = locatePieceIdx(t);
    pos[0] = 0.0; pos[1] = 0.0;
    double tn = 1.0;
    for (int k = 0; k <= 5; k++) {
      pos[0] += tn * coef[(6 * i + k) * 2];
      pos[1] += tn * coef[(6 * i + k) * 2 + 1];
      tn *= t;
    }
  }
  void getVel(double t, double vel[2]) const {
    int i = locatePieceIdx(t);
    vel[0] = 0.0; vel[1] = 0.0;
    double tn = 1.0;
    int n = 1;
    for (int k = 1; k <= 5; k++) {
      vel[0] += n * tn * coef[(6 * i + k) * 2];
      vel[1] += n * tn * coef[(6 * i + k) * 2 + 1];
      tn *= t;
      n++;
    }
  }
};

// =====================================================================================
// L1-L2  L-BFGS with Lewis-Overton line search (ALORE-modified)        lbf:276-390, 440-751
// =====================================================================================
enum {
  LBFGS_CONVERGENCE = 0, LBFGS_STOP, LBFGS_CANCELED,
  LBFGSERR_UNKNOWNERROR = -1024, LBFGSERR_INVALID_N, LBFGSERR_INVALID_MEMSIZE, LBFGSERR_INVALID_GEPSILON,
  LBFGSERR_INVALID_TESTPERIOD, LBFGSERR_INVALID_DELTA, LBFGSERR_INVALID_MINSTEP, LBFGSERR_INVALID_MAXSTEP,
  LBFGSERR_INVALID_FDECCOEFF, LBFGSERR_INVALID_SCURVCOEFF, LBFGSERR_INVALID_MACHINEPREC,
  LBFGSERR_INVALID_MAXLINESEARCH, LBFGSERR_INVALID_FUNCVAL, LBFGSERR_MINIMUMSTEP, LBFGSERR_MAXIMUMSTEP,
  LBFGSERR_MAXIMUMLINESEARCH, LBFGSERR_MAXIMUMITERATION, LBFGSERR_WIDTHTOOSMALL,
  LBFGSERR_INVALIDPARAMETERS, LBFGSERR_INCREASEGRADIENT,
};

using Vec = std::vector<double>;
// Summation order of VectorXd::dot / norm / squaredNorm: NOT defined by the reference's source (Eigen 3.3
// vectorises it with ISA-dependent packet partial sums).  The oracle fixes it to 32 strided partial sums
// (element i goes to partial i % 32, in index order) combined by a xor-butterfly (16, 8, 4, 2, 1) — a
// deterministic order that a 32-lane warp reproduces exactly.
inline double vdot(const double* a, const double* b, int n) {
  double p[32], q[32];
  for (int l = 0; l < 32; l++) p[l] = 0.0;
  for (int i = 0; i < n; i++) p[i & 31] += a[i] * b[i];
  for (int o = 16; o > 0; o >>= 1) {
    for (int l = 0; l < 32; l++) q[l] = p[l] + p[l ^ o];
    for (int l = 0; l < 32; l++) p[l] = q[l];
  }
  return p[0];
}
inline double vnorm(const double* a, int n) { return std::sqrt(vdot(a, a, n)); }
inline double vabsmax(const double* a, int n) { double m = std::fabs(a[0]); for (int i = 1; i < n; i++) m = std::max(m, std::fabs(a[i])); return m; }

template <typename Eval>
int line_search_lewisoverton(Vec& x, double& f, Vec& g, double& stp, const Vec& s, const Vec& xp, const Vec& gp,
                             double stpmin, double stpmax, Eval& eval, const alore_lbfgs_params_t& param) {
  const int n = (int)x.size();
  int count = 0;
  bool brackt = false, touched = false;
  double finit, dginit, dgtest, dstest;
  double mu = 0.0, nu = stpmax;
  if (!(stp > 0.0)) return LBFGSERR_INVALIDPARAMETERS;
  dginit = vdot(gp.data(), s.data(), n);
  if (0.0 < dginit) return LBFGSERR_INCREASEGRADIENT;
  finit = f;
  dgtest = param.f_dec_coeff * dginit;
  dstest = param.s_curv_coeff * dginit;
  while (true) {
    for (int i = 0; i < n; i++) x[i] = xp[i] + stp * s[i];
    f = eval(x, g);
    ++count;
    if (std::isinf(f) || std::isnan(f)) return LBFGSERR_INVALID_FUNCVAL;
    if (param.past > 0 && std::fabs(finit - f) / (std::fabs(finit) + 1.0) < param.delta / param.past) return count;  // lbf:326-329
    if (f > finit + stp * dgtest) {
      nu = stp;
      brackt = true;
    } else {
      if (vdot(g.data(), s.data(), n) < dstest) mu = stp;
      else return count;
    }
    if (param.max_linesearch <= count) return LBFGSERR_MAXIMUMLINESEARCH;
    if (brackt && (nu - mu) < param.machine_prec * nu) return LBFGSERR_WIDTHTOOSMALL;
    if (brackt) stp = 0.5 * (mu + nu);
    else stp *= 2.0;
    if (stp < stpmin) return LBFGSERR_MINIMUMSTEP;
    if (stp > stpmax) {
      if (touched) return LBFGSERR_MAXIMUMSTEP;
      touched = true;
      stp = stpmax;
    }
  }
}

// eval(x, g) -> f.  proc_stepbound is always NULL in the reference's calls; proc_progress is
// NULL (stage A) or earlyExit which always returns 0 (opt:593-628), so both are omitted.
template <typename Eval>
int lbfgs_optimize(Vec& x, double& f, Eval&& eval, const alore_lbfgs_params_t& param, int* iters_out = nullptr) {
  int ret, i, j, k, ls, end, bound;
  double step, step_min, step_max, fx, ys, yy;
  double gnorm_inf, xnorm_inf, beta, rate, cau;
  const int n = (int)x.size();
  const int m = param.mem_size;
  if (n <= 0) return LBFGSERR_INVALID_N;
  if (m <= 0) return LBFGSERR_INVALID_MEMSIZE;
  if (param.g_epsilon < 0.0) return LBFGSERR_INVALID_GEPSILON;
  if (param.past < 0) return LBFGSERR_INVALID_TESTPERIOD;
  if (param.delta < 0.0) return LBFGSERR_INVALID_DELTA;
  if (param.min_step < 0.0) return LBFGSERR_INVALID_MINSTEP;
  if (param.max_step < param.min_step) return LBFGSERR_INVALID_MAXSTEP;
  if (!(param.f_dec_coeff > 0.0 && param.f_dec_coeff < 1.0)) return LBFGSERR_INVALID_FDECCOEFF;
  if (!(param.s_curv_coeff < 1.0 && param.s_curv_coeff > param.f_dec_coeff)) return LBFGSERR_INVALID_SCURVCOEFF;
  if (!(param.machine_prec > 0.0)) return LBFGSERR_INVALID_MACHINEPREC;
  if (param.max_linesearch <= 0) return LBFGSERR_INVALID_MAXLINESEARCH;

  Vec xp(n), g(n), gp(n), d(n), pf(std::max(1, param.past));
  Vec lm_alpha(m, 0.0), lm_s(size_t(n) * m, 0.0), lm_y(size_t(n) * m, 0.0), lm_ys(m, 0.0);  // column j at [j*n, j*n+n)

  fx = eval(x, g);
  pf[0] = fx;
  for (i = 0; i < n; i++) d[i] = -g[i];
  gnorm_inf = vabsmax(g.data(), n);
  xnorm_inf = vabsmax(x.data(), n);
  k = 0;
  if (gnorm_inf / std::max(1.0, xnorm_inf) < param.g_epsilon) {
    ret = LBFGS_CONVERGENCE;
  } else {
    step = 1.0 / vnorm(d.data(), n);
    k = 1;
    end = 0;
    bound = 0;
    while (true) {
      xp = x;
      gp = g;
      step_min = param.min_step;
      step_max = param.max_step;
      ls = line_search_lewisoverton(x, fx, g, step, d, xp, gp, step_min, step_max, eval, param);
      if (ls < 0) {
        x = xp;
        g = gp;
        ret = ls;
        break;
      }
      gnorm_inf = vabsmax(g.data(), n);
      xnorm_inf = vabsmax(x.data(), n);
      if (gnorm_inf / std::max(1.0, xnorm_inf) < param.g_epsilon) { ret = LBFGS_CONVERGENCE; break; }
      if (0 < param.past) {
        if (param.past <= k) {
          rate = std::fabs(pf[k % param.past] - fx) / std::max(1.0, std::fabs(fx));
          if (rate < param.delta) { ret = LBFGS_STOP; break; }
        }
        pf[k % param.past] = fx;
      }
      if (param.max_iterations != 0 && param.max_iterations <= k) { ret = LBFGSERR_MAXIMUMITERATION; break; }
      ++k;
      double* sc = &lm_s[size_t(end) * n];
      double* yc = &lm_y[size_t(end) * n];
      for (i = 0; i < n; i++) { sc[i] = x[i] - xp[i]; yc[i] = g[i] - gp[i]; }
      ys = vdot(yc, sc, n);
      yy = vdot(yc, yc, n);
      lm_ys[end] = ys;
      for (i = 0; i < n; i++) d[i] = -g[i];
      cau = vdot(sc, sc, n) * vnorm(gp.data(), n) * param.cautious_factor;
      if (ys > cau) {
        ++bound;
        bound = m < bound ? m : bound;
        end = (end + 1) % m;
        j = end;
        for (i = 0; i < bound; ++i) {
          j = (j + m - 1) % m;
          lm_alpha[j] = vdot(&lm_s[size_t(j) * n], d.data(), n) / lm_ys[j];
          const double c = -lm_alpha[j];
          const double* yj = &lm_y[size_t(j) * n];
          for (int t = 0; t < n; t++) d[t] += c * yj[t];
        }
        {
          const double c = ys / yy;
          for (int t = 0; t < n; t++) d[t] *= c;
        }
        for (i = 0; i < bound; ++i) {
          beta = vdot(&lm_y[size_t(j) * n], d.data(), n) / lm_ys[j];
          const double c = lm_alpha[j] - beta;
          const double* sj = &lm_s[size_t(j) * n];
          for (int t = 0; t < n; t++) d[t] += c * sj[t];
          j = (j + 1) % m;
        }
      }
      step = 1.0;
    }
  }
  f = fx;
  if (iters_out) *iters_out = k;
  return ret;
}

// =====================================================================================
// FlatTrajData                                                traj_representation.h:46-76
// =====================================================================================
struct FlatTrajData {
  std::vector<std::array<double, 3>> UnOccupied_traj_pts;   // yaw, s, t
  double UnOccupied_initT = 0.0;
  std::vector<std::array<double, 3>> UnOccupied_positions;  // x, y, yaw
  double start_state[2][3]{};   // rows: yaw, s; cols: P,V,A
  double final_state[2][3]{};
  double start_state_XYTheta[3]{};
  double final_state_XYTheta[3]{};
  bool if_cut = false;
};

// =====================================================================================
// P1-P4, O1-O3  MSPlanner                                                 opt:169-1106, 1272-1591
// =====================================================================================
struct MSPlanner {
  alore_params_t p{};
  const SdfMap* map_ = nullptr;

  // mutable planner state (members of the reference class, optimizer.h:110-175)
  double time_weight = 0.0;       // penaltyWt.time_weight (scaled by 0.75 on collision replans)
  double safeDis = 0.0;
  int TrajNum = 0;
  bool ifCutTraj_ = false;
  int unOccupied_traj_num_ = -1;
  Vec pieceTime, Innerpoints /*2 x (N-1) col-major*/, finalInnerpoints, finalpieceTime;
  double iniState[2][3]{}, finState[2][3]{};
  double iniStateXYTheta[3]{}, finStateXYTheta[3]{};
  std::vector<std::array<double, 3>> inner_init_positions;
  MincoS3NU Minco;
  int iter_num_ = 0;
  long total_evals = 0;
  Vec gradByPoints, gradByTimes, partialGradByCoeffs, partialGradByTimes;
  double gradByTailStateS[2]{};
  double FinalIntegralXYError[2]{};
  double EqualLambda[2]{}, EqualRho[2]{};
  int SamNumEachPart = 0, sparseResolution_ = 0, sparseResolution_6_ = 0;
  Vec IntegralChainCoeff;
  Traj optimizer_traj_, final_traj_;
  // diagnostics (not in the reference)
  int last_status = 0, last_alm_iters = 0, replans = 0;
  double last_cost = 0.0;

  void init(const alore_params_t& prm, const SdfMap* map) {
    p = prm;
    map_ = map;
    time_weight = p.pw_time;
    safeDis = p.safeDis;
    sparseResolution_ = p.sparseResolution;
    SamNumEachPart = 2 * sparseResolution_;        // opt:141
    sparseResolution_6_ = sparseResolution_ * 6;    // opt:142
    IntegralChainCoeff.assign(SamNumEachPart + 1, 0.0);  // opt:143-147
    for (int i = 0; i < sparseResolution_; i++) {
      IntegralChainCoeff[2 * i] += 1.0;
      IntegralChainCoeff[2 * i + 1] += 4.0;
      IntegralChainCoeff[2 * i + 2] += 1.0;
    }
  }

  // opt:573-591
  static void RealT2VirtualT(const Vec& RT, double* VT) {
    for (size_t i = 0; i < RT.size(); ++i)
      VT[i] = RT[i] > 1.0 ? (std::sqrt(2.0 * RT[i] - 1.0) - 1.0) : (1.0 - std::sqrt(2.0 / RT[i] - 1.0));
  }
  static void VirtualT2RealT(const double* VT, int n, Vec& RT) {
    RT.resize(n);
    for (int i = 0; i < n; ++i)
      RT[i] = VT[i] > 0.0 ? ((0.5 * VT[i] + 1.0) * VT[i] + 1.0) : 1.0 / ((0.5 * VT[i] - 1.0) * VT[i] + 1.0);
  }
  // opt:1088-1106
  static void backwardGradT(const double* tau, const Vec& gradT, double* gradTau) {
    for (size_t i = 0; i < gradT.size(); i++) {
      double gradrt2vt;
      if (tau[i] > 0) gradrt2vt = tau[i] + 1.0;
      else {
        double denSqrt = (0.5 * tau[i] - 1.0) * tau[i] + 1.0;
        gradrt2vt = (1.0 - tau[i]) / (denSqrt * denSqrt);
      }
      gradTau[i] = gradT[i] * gradrt2vt;
    }
  }
  // opt:1069-1086
  void positiveSmoothedL1(double x, double& f, double& df) const {
    const double pe = p.smoothEps;
    const double half = 0.5 * pe;
    const double f3c = 1.0 / (pe * pe);
    const double f4c = -0.5 * f3c / pe;
    const double d2c = 3.0 * f3c;
    const double d3c = 4.0 * f4c;
    if (x < pe) {
      f = (f4c * x + f3c) * x * x * x;
      df = (d3c * x + d2c) * x * x;
    } else {
      f = x - half;
      df = 1.0;
    }
  }

  // Per-sample chain-rule columns shared by both penalty functionals (opt:787-826 / 966-994).
  struct ChainCols {
    std::vector<double> XGradCS, XGradCTheta, XGradT, YGradCS, YGradCTheta, YGradT;  // 6 x (S+1) col-major / (S+1)
    void resize(int S1) {
      XGradCS.assign(size_t(6) * S1, 0.0); XGradCTheta = XGradCS; YGradCS = XGradCS; YGradCTheta = XGradCS;
      XGradT.assign(S1, 0.0); YGradT = XGradT;
    }
  };

  // Integral contributions + chain-rule columns for one sample; `wint` is 1 (even) or 4 (odd).
  void sampleIntegral(int j, const double beta0[6], const double beta1[6], const double sigma[2], const double dsigma[2],
                      const double ddsigma[2], double cosyaw, double sinyaw, double CoeffIntegral, double IntegralAlpha,
                      Vec& IntegralX, Vec& IntegralY, ChainCols& cc) const {
    const bool even = (j % 2 == 0);
    const double icr = p.ICR[2];
    (void)sigma;
    if (p.if_standard_diff) {
      if (even) {
        if (j != 0) {
          IntegralX[j / 2 - 1] += CoeffIntegral * dsigma[1] * cosyaw;
          IntegralY[j / 2 - 1] += CoeffIntegral * dsigma[1] * sinyaw;
        }
        if (j != SamNumEachPart) {
          IntegralX[j / 2] += CoeffIntegral * dsigma[1] * cosyaw;
          IntegralY[j / 2] += CoeffIntegral * dsigma[1] * sinyaw;
        }
      } else {
        IntegralX[j / 2] += 4 * CoeffIntegral * dsigma[1] * cosyaw;
        IntegralY[j / 2] += 4 * CoeffIntegral * dsigma[1] * sinyaw;
      }
      for (int r = 0; r < 6; r++) {
        cc.XGradCS[6 * j + r] = beta1[r] * cosyaw;
        cc.XGradCTheta[6 * j + r] = -dsigma[1] * beta0[r] * sinyaw;
        cc.YGradCS[6 * j + r] = beta1[r] * sinyaw;
        cc.YGradCTheta[6 * j + r] = dsigma[1] * beta0[r] * cosyaw;
      }
      cc.XGradT[j] = (ddsigma[1] * cosyaw - dsigma[1] * dsigma[0] * sinyaw) * IntegralAlpha * CoeffIntegral + dsigma[1] * cosyaw / sparseResolution_6_;
      cc.YGradT[j] = (ddsigma[1] * sinyaw + dsigma[1] * dsigma[0] * cosyaw) * IntegralAlpha * CoeffIntegral + dsigma[1] * sinyaw / sparseResolution_6_;
    } else {
      const double ix = (dsigma[1] * cosyaw + dsigma[0] * icr * sinyaw);
      const double iy = (dsigma[1] * sinyaw - dsigma[0] * icr * cosyaw);
      if (even) {
        if (j != 0) {
          IntegralX[j / 2 - 1] += CoeffIntegral * ix;
          IntegralY[j / 2 - 1] += CoeffIntegral * iy;
        }
        if (j != SamNumEachPart) {
          IntegralX[j / 2] += CoeffIntegral * ix;
          IntegralY[j / 2] += CoeffIntegral * iy;
        }
      } else {
        IntegralX[j / 2] += 4 * CoeffIntegral * ix;
        IntegralY[j / 2] += 4 * CoeffIntegral * iy;
      }
      for (int r = 0; r < 6; r++) {
        cc.XGradCS[6 * j + r] = beta1[r] * cosyaw;
        cc.XGradCTheta[6 * j + r] = beta0[r] * (-dsigma[1] * sinyaw + dsigma[0] * icr * cosyaw) + beta1[r] * sinyaw * icr;
        cc.YGradCS[6 * j + r] = beta1[r] * sinyaw;
        cc.YGradCTheta[6 * j + r] = beta0[r] * (dsigma[1] * cosyaw - dsigma[0] * icr * sinyaw) - beta1[r] * cosyaw * icr;
      }
      cc.XGradT[j] = (ddsigma[1] * cosyaw - dsigma[1] * dsigma[0] * sinyaw + ddsigma[0] * icr * sinyaw + dsigma[0] * dsigma[0] * icr * cosyaw) * IntegralAlpha * CoeffIntegral
                     + (dsigma[1] * cosyaw + dsigma[0] * icr * sinyaw) / sparseResolution_6_;
      cc.YGradT[j] = (ddsigma[1] * sinyaw + dsigma[1] * dsigma[0] * cosyaw - ddsigma[0] * icr * cosyaw + dsigma[0] * dsigma[0] * icr * sinyaw) * IntegralAlpha * CoeffIntegral
                     + (dsigma[1] * sinyaw - dsigma[0] * icr * cosyaw) / sparseResolution_6_;
    }
  }

  static inline void polyBasis(double s1, double b0[6], double b1[6], double b2[6], double b3[6]) {
    double s2 = s1 * s1, s3 = s2 * s1, s4 = s2 * s2, s5 = s3 * s2;
    b0[0] = 1.0; b0[1] = s1; b0[2] = s2; b0[3] = s3; b0[4] = s4; b0[5] = s5;
    b1[0] = 0.0; b1[1] = 1.0; b1[2] = 2.0 * s1; b1[3] = 3.0 * s2; b1[4] = 4.0 * s3; b1[5] = 5.0 * s4;
    b2[0] = 0.0; b2[1] = 0.0; b2[2] = 2.0; b2[3] = 6.0 * s1; b2[4] = 12.0 * s2; b2[5] = 20.0 * s3;
    b3[0] = 0.0; b3[1] = 0.0; b3[2] = 0.0; b3[3] = 6.0; b3[4] = 24.0 * s1; b3[5] = 60.0 * s2;
  }
  // c.transpose() * beta for the 6x2 block of piece i.
  inline void cTb(int i, const double beta[6], double out[2]) const {
    for (int d = 0; d < 2; d++) {
      double s = 0.0;
      for (int r = 0; r < 6; r++) s += Minco.B(6 * i + r, d) * beta[r];
      out[d] = s;
    }
  }
  // Final chain push shared by both functionals (opt:1054-1066 / 1583-1590).
  void pushChain(const std::vector<ChainCols>& vec, const Vec& chainX, const Vec& chainY, const Vec& CI) {
    const int S1 = SamNumEachPart + 1;
    for (int i = 0; i < TrajNum; i++) {
      Vec CoeffX(S1), CoeffY(S1);
      for (int j = 0; j < S1; j++) {
        CoeffX[j] = chainX[size_t(i) * S1 + j] * IntegralChainCoeff[j];
        CoeffY[j] = chainY[size_t(i) * S1 + j] * IntegralChainCoeff[j];
      }
      const ChainCols& c = vec[i];
      double a1[6] = {0}, a2[6] = {0}, a3[6] = {0}, a4[6] = {0};
      for (int j = 0; j < S1; j++)
        for (int r = 0; r < 6; r++) {
          a1[r] += (c.XGradCS[6 * j + r] * CI[i]) * CoeffX[j];
          a2[r] += (c.XGradCTheta[6 * j + r] * CI[i]) * CoeffX[j];
          a3[r] += (c.YGradCS[6 * j + r] * CI[i]) * CoeffY[j];
          a4[r] += (c.YGradCTheta[6 * j + r] * CI[i]) * CoeffY[j];
        }
      for (int r = 0; r < 6; r++) {
        partialGradByCoeffs[2 * (6 * i + r) + 1] += a1[r];
        partialGradByCoeffs[2 * (6 * i + r) + 0] += a2[r];
        partialGradByCoeffs[2 * (6 * i + r) + 1] += a3[r];
        partialGradByCoeffs[2 * (6 * i + r) + 0] += a4[r];
      }
      double sx = 0.0, sy = 0.0;
      for (int j = 0; j < S1; j++) sx += c.XGradT[j] * CoeffX[j];
      for (int j = 0; j < S1; j++) sy += c.YGradT[j] * CoeffY[j];
      partialGradByTimes[i] += sx;
      partialGradByTimes[i] += sy;
    }
  }

  // opt:694-1067.  Uses Minco's current coefficients, pieceTime, EqualLambda/Rho, safeDis.
  void attachPenaltyFunctional(double& cost) {
    const double ini_x = iniStateXYTheta[0], ini_y = iniStateXYTheta[1];
    double beta0[6], beta1[6], beta2[6], beta3[6];
    double s1;
    double sigma[2], dsigma[2], ddsigma[2], dddsigma[2];
    double IntegralAlpha, Alpha, omg, omgstep;
    double violaAcc, violaAlp, violaPos, violaMom, violaCenAcc;
    double violaAccPena, violaAlpPena, violaPosPena, violaMomPena, violaCenAccPena;
    double violaAccPenaD, violaAlpPenaD, violaPosPenaD, violaMomPenaD, violaCenAccPenaD;
    double gradViolaAT, gradViolaDOT, gradViolaPt, gradViolaMt, gradViolaCAt;
    double violaVel, violaVelPena, violaVelPenaD, violaOmega, violaOmegaPena, violaOmegaPenaD;
    const int S1 = SamNumEachPart + 1;
    std::vector<Vec> VecIntegralX, VecIntegralY;
    std::vector<std::array<double, 2>> VecTrajFinalXY;
    VecTrajFinalXY.push_back({ini_x, ini_y});
    std::vector<ChainCols> VecCols;
    Vec VecCI;
    ChainCols cc;
    cc.resize(S1);
    Vec IntegralX(sparseResolution_), IntegralY(sparseResolution_);
    Vec VecCoeffChainX(size_t(TrajNum) * S1, 0.0), VecCoeffChainY(size_t(TrajNum) * S1, 0.0);
    double CurrentPointXY[2] = {ini_x, ini_y};
    const double max_vel_ = p.max_vel, min_vel_ = p.min_vel, max_acc_ = p.max_acc, max_omega_ = p.max_omega,
                 max_domega_ = p.max_domega, max_cen = p.max_centripetal_acc;

    for (int i = 0; i < TrajNum; i++) {
      double step = pieceTime[i] / sparseResolution_;
      double halfstep = step / 2.0;
      double CoeffIntegral = pieceTime[i] / sparseResolution_6_;
      std::fill(IntegralX.begin(), IntegralX.end(), 0.0);
      std::fill(IntegralY.begin(), IntegralY.end(), 0.0);
      s1 = 0.0;
      for (int j = 0; j <= SamNumEachPart; j++) {
        if (j % 2 == 0) {
          polyBasis(s1, beta0, beta1, beta2, beta3);
          s1 += halfstep;
          IntegralAlpha = 1.0 / SamNumEachPart * j;
          Alpha = 1.0 / sparseResolution_ * (double(j) / 2);
          omg = (j == 0 || j == SamNumEachPart) ? 0.5 : 1;
          omgstep = omg * step;
          cTb(i, beta0, sigma); cTb(i, beta1, dsigma); cTb(i, beta2, ddsigma); cTb(i, beta3, dddsigma);
          double gradBeta[3][2] = {{0, 0}, {0, 0}, {0, 0}};
          double cosyaw = ocos(sigma[0]), sinyaw = osin(sigma[0]);
          sampleIntegral(j, beta0, beta1, sigma, dsigma, ddsigma, cosyaw, sinyaw, CoeffIntegral, IntegralAlpha, IntegralX, IntegralY, cc);

          violaAcc = ddsigma[1] * ddsigma[1] - max_acc_ * max_acc_;
          violaAlp = ddsigma[0] * ddsigma[0] - max_domega_ * max_domega_;
          if (violaAcc > 0) {
            positiveSmoothedL1(violaAcc, violaAccPena, violaAccPenaD);
            gradViolaAT = 2.0 * Alpha * ddsigma[1] * dddsigma[1];
            gradBeta[2][1] += omgstep * p.pw_acc * violaAccPenaD * 2.0 * ddsigma[1];
            partialGradByTimes[i] += omg * p.pw_acc * (violaAccPenaD * gradViolaAT * step + violaAccPena / sparseResolution_);
            cost += omgstep * p.pw_acc * violaAccPena;
          }
          if (violaAlp > 0) {
            positiveSmoothedL1(violaAlp, violaAlpPena, violaAlpPenaD);
            gradViolaDOT = 2.0 * Alpha * ddsigma[0] * dddsigma[0];
            gradBeta[2][0] += omgstep * p.pw_domega * violaAlpPenaD * 2.0 * ddsigma[0];
            partialGradByTimes[i] += omg * p.pw_domega * (violaAlpPenaD * gradViolaDOT * step + violaAlpPena / sparseResolution_);
            cost += omgstep * p.pw_domega * violaAlpPena;
          }
          if (p.if_directly_constrain_v_omega) {
            violaVel = dsigma[1] * dsigma[1] - max_vel_ * max_vel_;
            if (violaVel > 0) {
              positiveSmoothedL1(violaVel, violaVelPena, violaVelPenaD);
              gradViolaPt = 2.0 * Alpha * dsigma[1] * ddsigma[1];
              gradBeta[1][1] += omgstep * p.pw_moment * violaVelPenaD * 2.0 * dsigma[1];
              partialGradByTimes[i] += omg * p.pw_moment * (violaVelPenaD * gradViolaPt * step + violaVelPena / sparseResolution_);
              cost += omgstep * p.pw_moment * violaVelPena;
            }
            violaOmega = dsigma[0] * dsigma[0] - max_omega_ * max_omega_;
            if (violaOmega > 0) {
              positiveSmoothedL1(violaOmega, violaOmegaPena, violaOmegaPenaD);
              gradViolaPt = 2.0 * Alpha * dsigma[0] * ddsigma[0];
              gradBeta[1][0] += omgstep * p.pw_moment * violaOmegaPenaD * 2.0 * dsigma[0];
              partialGradByTimes[i] += omg * p.pw_moment * (violaOmegaPenaD * gradViolaPt * step + violaOmegaPena / sparseResolution_);
              cost += omgstep * p.pw_moment * violaOmegaPena;
            }
          } else {
            for (int omg_sym = -1; omg_sym <= 1; omg_sym += 2) {
              violaMom = omg_sym * max_vel_ * dsigma[0] + max_omega_ * dsigma[1] - max_vel_ * max_omega_;
              if (violaMom > 0) {
                positiveSmoothedL1(violaMom, violaMomPena, violaMomPenaD);
                gradViolaMt = Alpha * (omg_sym * max_vel_ * ddsigma[0] + max_omega_ * ddsigma[1]);
                gradBeta[1][0] += omgstep * p.pw_moment * violaMomPenaD * omg_sym * max_vel_;
                gradBeta[1][1] += omgstep * p.pw_moment * violaMomPenaD * max_omega_;
                partialGradByTimes[i] += omg * p.pw_moment * (violaMomPenaD * gradViolaMt * step + violaMomPena / sparseResolution_);
                cost += omgstep * p.pw_moment * violaMomPena;
              }
            }
            for (int omg_sym = -1; omg_sym <= 1; omg_sym += 2) {
              violaMom = omg_sym * -min_vel_ * dsigma[0] - max_omega_ * dsigma[1] + min_vel_ * max_omega_;
              if (violaMom > 0) {
                positiveSmoothedL1(violaMom, violaMomPena, violaMomPenaD);
                gradViolaMt = Alpha * (omg_sym * -min_vel_ * ddsigma[0] - max_omega_ * ddsigma[1]);
                gradBeta[1][0] += omgstep * p.pw_moment * violaMomPenaD * omg_sym * -min_vel_;
                gradBeta[1][1] -= omgstep * p.pw_moment * violaMomPenaD * max_omega_;
                partialGradByTimes[i] += omg * p.pw_moment * (violaMomPenaD * gradViolaMt * step + violaMomPena / sparseResolution_);
                cost += omgstep * p.pw_moment * violaMomPena;
              }
            }
          }
          violaCenAcc = dsigma[0] * dsigma[0] * dsigma[1] * dsigma[1] - max_cen * max_cen;
          if (violaCenAcc > 0) {
            positiveSmoothedL1(violaCenAcc, violaCenAccPena, violaCenAccPenaD);
            gradViolaCAt = 2.0 * Alpha * (dsigma[0] * dsigma[1] * dsigma[1] * ddsigma[0] + dsigma[1] * dsigma[0] * dsigma[0] * ddsigma[1]);
            gradBeta[1][0] += omgstep * p.pw_cen_acc * violaCenAccPenaD * (2 * dsigma[0] * dsigma[1] * dsigma[1]);
            gradBeta[1][1] += omgstep * p.pw_cen_acc * violaCenAccPenaD * (2 * dsigma[0] * dsigma[0] * dsigma[1]);
            partialGradByTimes[i] += omg * p.pw_cen_acc * (violaCenAccPenaD * gradViolaCAt * step + violaCenAccPena / sparseResolution_);
            cost += omgstep * p.pw_cen_acc * violaCenAccPena;
          }

          // collision                                                         opt:912-947
          if (j != 0) {
            CurrentPointXY[0] += IntegralX[j / 2 - 1];
            CurrentPointXY[1] += IntegralY[j / 2 - 1];
          }
          bool if_coolision = false;
          double all_grad2Pos[2] = {0.0, 0.0};
          for (int c = 0; c < p.n_checkpoints; c++) {
            const double cpx = p.check_point[c][0], cpy = p.check_point[c][1];
            double bpt[2] = {CurrentPointXY[0] + (cosyaw * cpx + (-sinyaw) * cpy),
                             CurrentPointXY[1] + (sinyaw * cpx + cosyaw * cpy)};
            double gradESDF2d[2] = {0.0, 0.0};
            double sdf_value = map_->getDistWithGradBilinear(bpt, gradESDF2d, safeDis);
            violaPos = -sdf_value + safeDis;
            if (violaPos > 0.0) {
              if_coolision = true;
              positiveSmoothedL1(violaPos, violaPosPena, violaPosPenaD);
              const double sc = omgstep * p.pw_collision * violaPosPenaD;
              all_grad2Pos[0] -= sc * gradESDF2d[0];
              all_grad2Pos[1] -= sc * gradESDF2d[1];
              // help_L << -sin, -cos, cos, -sin
              const double L00 = -sinyaw, L01 = -cosyaw, L10 = cosyaw, L11 = -sinyaw;
              {
                const double sA = -Alpha * dsigma[0];
                const double r0 = (sA * gradESDF2d[0]) * L00 + (sA * gradESDF2d[1]) * L10;
                const double r1 = (sA * gradESDF2d[0]) * L01 + (sA * gradESDF2d[1]) * L11;
                gradViolaPt = r0 * cpx + r1 * cpy;
              }
              {
                const double r0 = (sc * gradESDF2d[0]) * L00 + (sc * gradESDF2d[1]) * L10;
                const double r1 = (sc * gradESDF2d[0]) * L01 + (sc * gradESDF2d[1]) * L11;
                gradBeta[0][0] -= r0 * cpx + r1 * cpy;
              }
              partialGradByTimes[i] += omg * p.pw_collision * (violaPosPenaD * gradViolaPt * step + violaPosPena / sparseResolution_);
              cost += omgstep * p.pw_collision * violaPosPena;
            }
          }
          if (if_coolision && !g_exact_chain_weights) {
            const int cnt = i * S1 + j + 1;
            for (int t = 0; t < cnt; t++) { VecCoeffChainX[t] += all_grad2Pos[0]; VecCoeffChainY[t] += all_grad2Pos[1]; }
          } else if (if_coolision) {
            // TEST HOOK (not the reference): exact Simpson weight of the colliding sample itself, see g_exact_chain_weights
            const int cnt = i * S1 + j;
            for (int t = 0; t < cnt; t++) { VecCoeffChainX[t] += all_grad2Pos[0]; VecCoeffChainY[t] += all_grad2Pos[1]; }
            const double ow = (j == 0) ? 0.0 : ((j == SamNumEachPart) ? 1.0 : 0.5);
            VecCoeffChainX[cnt] += ow * all_grad2Pos[0];
            VecCoeffChainY[cnt] += ow * all_grad2Pos[1];
          }
          for (int r = 0; r < 6; r++)
            for (int d = 0; d < 2; d++)
              partialGradByCoeffs[2 * (6 * i + r) + d] += beta0[r] * gradBeta[0][d] + beta1[r] * gradBeta[1][d] + beta2[r] * gradBeta[2][d];
        } else {
          polyBasis(s1, beta0, beta1, beta2, beta3);
          s1 += halfstep;
          IntegralAlpha = 1.0 / SamNumEachPart * j;
          cTb(i, beta0, sigma); cTb(i, beta1, dsigma); cTb(i, beta2, ddsigma);
          double cosyaw = ocos(sigma[0]), sinyaw = osin(sigma[0]);
          sampleIntegral(j, beta0, beta1, sigma, dsigma, ddsigma, cosyaw, sinyaw, CoeffIntegral, IntegralAlpha, IntegralX, IntegralY, cc);
        }
      }
      // mean-time penalty is dead code: unOccupied_traj_num_ = -1 (opt:225, 999)
      VecIntegralX.push_back(IntegralX);
      VecIntegralY.push_back(IntegralY);
      double sx = 0.0, sy = 0.0;
      for (int t = 0; t < sparseResolution_; t++) { sx += IntegralX[t]; sy += IntegralY[t]; }
      VecTrajFinalXY.push_back({VecTrajFinalXY[i][0] + sx, VecTrajFinalXY[i][1] + sy});
      VecCols.push_back(cc);
      VecCI.push_back(CoeffIntegral);
    }
    // final position constraint                                               opt:1027-1037
    FinalIntegralXYError[0] = VecTrajFinalXY.back()[0] - finStateXYTheta[0];
    FinalIntegralXYError[1] = VecTrajFinalXY.back()[1] - finStateXYTheta[1];
    {
      const double ax = FinalIntegralXYError[0] + EqualLambda[0] / EqualRho[0];
      const double ay = FinalIntegralXYError[1] + EqualLambda[1] / EqualRho[1];
      cost += 0.5 * (EqualRho[0] * (ax * ax) + EqualRho[1] * (ay * ay));
      const double cx = EqualRho[0] * ax, cy = EqualRho[1] * ay;
      for (auto& v : VecCoeffChainX) v += cx;
      for (auto& v : VecCoeffChainY) v += cy;
    }
    pushChain(VecCols, VecCoeffChainX, VecCoeffChainY, VecCI);
  }

  // opt:1319-1591
  void attachPenaltyFunctionalPath(double& cost) {
    const double ini_x = iniStateXYTheta[0], ini_y = iniStateXYTheta[1];
    double beta0[6], beta1[6], beta2[6], beta3[6];
    double s1;
    double sigma[2], dsigma[2], ddsigma[2], dddsigma[2];
    const int S1 = SamNumEachPart + 1;
    double IntegralAlpha, omg;
    double violaPos, violaMom, violaMomPena, violaMomPenaD;
    std::vector<std::array<double, 2>> VecTrajFinalXY(TrajNum + 1);
    VecTrajFinalXY[0] = {ini_x, ini_y};
    std::vector<ChainCols> VecCols(TrajNum);
    Vec VecCI(TrajNum);
    Vec VecCoeffChainX(size_t(TrajNum) * S1, 0.0), VecCoeffChainY(size_t(TrajNum) * S1, 0.0);
    const double max_vel_ = p.max_vel, min_vel_ = p.min_vel, max_acc_ = p.max_acc, max_omega_ = p.max_omega, max_domega_ = p.max_domega;

    for (int i = 0; i < TrajNum; i++) {
      double step = pieceTime[i] / sparseResolution_;
      double halfstep = step / 2;
      double CoeffIntegral = pieceTime[i] / sparseResolution_ / 6;
      ChainCols cc;
      cc.resize(S1);
      Vec IntegralX(sparseResolution_, 0.0), IntegralY(sparseResolution_, 0.0);
      s1 = 0.0;
      for (int j = 0; j <= SamNumEachPart; j++) {
        if (j % 2 == 0) {
          polyBasis(s1, beta0, beta1, beta2, beta3);
          s1 += halfstep;
          IntegralAlpha = 1.0 / SamNumEachPart * j;
          omg = (j == 0 || j == SamNumEachPart) ? 0.5 : 1;
          cTb(i, beta0, sigma); cTb(i, beta1, dsigma); cTb(i, beta2, ddsigma); cTb(i, beta3, dddsigma);
          double cosyaw = ocos(sigma[0]), sinyaw = osin(sigma[0]);
          sampleIntegral(j, beta0, beta1, sigma, dsigma, ddsigma, cosyaw, sinyaw, CoeffIntegral, IntegralAlpha, IntegralX, IntegralY, cc);
          double gradViolaMt;
          double Alpha = 1.0 / sparseResolution_ * (double(j) / 2);
          double gradBeta[3][2] = {{0, 0}, {0, 0}, {0, 0}};
          for (int omg_sym = -1; omg_sym <= 1; omg_sym += 2) {
            violaMom = omg_sym * max_vel_ * dsigma[0] + max_omega_ * dsigma[1] - max_vel_ * max_omega_;
            if (violaMom > 0) {
              positiveSmoothedL1(violaMom, violaMomPena, violaMomPenaD);
              gradViolaMt = Alpha * (omg_sym * max_vel_ * ddsigma[0] + max_omega_ * ddsigma[1]);
              gradBeta[1][0] += omg * step * p.ppw_moment * violaMomPenaD * omg_sym * max_vel_;
              gradBeta[1][1] += omg * step * p.ppw_moment * violaMomPenaD * max_omega_;
              partialGradByTimes[i] += omg * p.ppw_moment * (violaMomPenaD * gradViolaMt * step + violaMomPena / sparseResolution_);
              cost += omg * step * p.ppw_moment * violaMomPena;
            }
          }
          for (int omg_sym = -1; omg_sym <= 1; omg_sym += 2) {
            violaMom = omg_sym * -min_vel_ * dsigma[0] - max_omega_ * dsigma[1] + min_vel_ * max_omega_;
            if (violaMom > 0) {
              positiveSmoothedL1(violaMom, violaMomPena, violaMomPenaD);
              gradViolaMt = Alpha * (omg_sym * -min_vel_ * ddsigma[0] - max_omega_ * ddsigma[1]);
              gradBeta[1][0] += omg * step * p.ppw_moment * violaMomPenaD * omg_sym * -min_vel_;
              gradBeta[1][1] -= omg * step * p.ppw_moment * violaMomPenaD * max_omega_;
              partialGradByTimes[i] += omg * p.ppw_moment * (violaMomPenaD * gradViolaMt * step + violaMomPena / sparseResolution_);
              cost += omg * step * p.ppw_moment * violaMomPena;
            }
          }
          double violaAcc = ddsigma[1] * ddsigma[1] - max_acc_ * max_acc_;
          double violaAlp = ddsigma[0] * ddsigma[0] - max_domega_ * max_domega_;
          double violaAccPena, violaAccPenaD, violaAlpPena, violaAlpPenaD;
          if (violaAcc > 0) {
            positiveSmoothedL1(violaAcc, violaAccPena, violaAccPenaD);
            double gradViolaAT = 2.0 * Alpha * ddsigma[1] * dddsigma[1];
            gradBeta[2][1] += omg * step * p.ppw_acc * violaAccPenaD * 2.0 * ddsigma[1];
            partialGradByTimes[i] += omg * p.ppw_acc * (violaAccPenaD * gradViolaAT * step + violaAccPena / sparseResolution_);
            cost += omg * step * p.ppw_acc * violaAccPena;
          }
          if (violaAlp > 0) {
            positiveSmoothedL1(violaAlp, violaAlpPena, violaAlpPenaD);
            double gradViolaDOT = 2.0 * Alpha * ddsigma[0] * dddsigma[0];
            gradBeta[2][0] += omg * step * p.ppw_domega * violaAlpPenaD * 2.0 * ddsigma[0];
            partialGradByTimes[i] += omg * p.ppw_domega * (violaAlpPenaD * gradViolaDOT * step + violaAlpPena / sparseResolution_);
            cost += omg * step * p.ppw_domega * violaAlpPena;
          }
          for (int r = 0; r < 6; r++)
            for (int d = 0; d < 2; d++)
              partialGradByCoeffs[2 * (6 * i + r) + d] += beta0[r] * gradBeta[0][d] + beta1[r] * gradBeta[1][d] + beta2[r] * gradBeta[2][d];
        } else {
          polyBasis(s1, beta0, beta1, beta2, beta3);
          s1 += halfstep;
          IntegralAlpha = 1.0 / SamNumEachPart * j;
          cTb(i, beta0, sigma); cTb(i, beta1, dsigma); cTb(i, beta2, ddsigma);
          double cosyaw = ocos(sigma[0]), sinyaw = osin(sigma[0]);
          sampleIntegral(j, beta0, beta1, sigma, dsigma, ddsigma, cosyaw, sinyaw, CoeffIntegral, IntegralAlpha, IntegralX, IntegralY, cc);
        }
      }
      double sx = 0.0, sy = 0.0;
      for (int t = 0; t < sparseResolution_; t++) { sx += IntegralX[t]; sy += IntegralY[t]; }
      VecTrajFinalXY[i + 1] = {VecTrajFinalXY[i][0] + sx, VecTrajFinalXY[i][1] + sy};
      VecCols[i] = cc;
      VecCI[i] = CoeffIntegral;
      // path point constraint                                                 opt:1566-1572
      const double ipx = VecTrajFinalXY[i + 1][0], ipy = VecTrajFinalXY[i + 1][1];
      const double dx = ipx - inner_init_positions[i][0], dy = ipy - inner_init_positions[i][1];
      violaPos = dx * dx + dy * dy;
      const int cnt = (i + 1) * S1;
      const double ax = p.ppw_bigpath_sdf * 2.0 * (ipx - inner_init_positions[i][0]);
      const double ay = p.ppw_bigpath_sdf * 2.0 * (ipy - inner_init_positions[i][1]);
      for (int t = 0; t < cnt; t++) { VecCoeffChainX[t] += ax; VecCoeffChainY[t] += ay; }
      cost += p.ppw_bigpath_sdf * violaPos;
    }
    pushChain(VecCols, VecCoeffChainX, VecCoeffChainY, VecCI);
  }

  // opt:631-692 (stage 1) and opt:1272-1317 (stage 0 = Path).  `inf` macro = 1>>30 = 0.
  double costFunction(int stage, const Vec& x, Vec& g) {
    if (vnorm(x.data(), (int)x.size()) > 1e4) return 0;  // traj_representation.h:21, g untouched
    iter_num_ += 1;
    total_evals += 1;
    if (stage == 1) std::fill(g.begin(), g.end(), 0.0);  // g.setZero() only in costFunctionCallback (opt:641)
    int offset = 0;
    const double* P = x.data();
    double* gradP = g.data();
    offset += 2 * (TrajNum - 1);
    double* gradTailS = g.data() + offset;
    finState[1][0] = x[offset];
    ++offset;
    for (int t = 0; t < 2 * (TrajNum - 1); t++) gradP[t] = 0.0;
    Innerpoints.assign(P, P + 2 * (TrajNum - 1));
    const double* t_ = x.data() + offset;
    double* gradt = g.data() + offset;
    VirtualT2RealT(t_, TrajNum, pieceTime);
    for (int t = 0; t < TrajNum; t++) gradt[t] = 0.0;
    double cost;
    Minco.setTConditions(finState);
    Minco.setParameters(Innerpoints.data(), pieceTime.data());
    Minco.getEnergy(cost);
    Minco.getEnergyPartialGradByCoeffs(partialGradByCoeffs);
    Minco.getEnergyPartialGradByTimes(partialGradByTimes);
    if (stage == 1) attachPenaltyFunctional(cost);
    else attachPenaltyFunctionalPath(cost);
    Minco.propogateArcYawLenghGrad(partialGradByCoeffs, partialGradByTimes, gradByPoints, gradByTimes, gradByTailStateS);
    double tsum = 0.0;
    for (int t = 0; t < TrajNum; t++) tsum += pieceTime[t];
    if (stage == 1) cost += time_weight * tsum;          // opt:678
    else cost += p.ppw_time * tsum;                       // opt:1308 (PathpenaltyWt for the cost ...)
    for (int t = 0; t < TrajNum; t++) gradByTimes[t] += time_weight * 1.0;  // opt:684 / 1312 (... penaltyWt for the gradient)
    *gradTailS = gradByTailStateS[1];
    for (int t = 0; t < 2 * (TrajNum - 1); t++) gradP[t] = gradByPoints[t];
    backwardGradT(t_, gradByTimes, gradt);
    return cost;
  }

  // opt:222-249
  void get_state(const FlatTrajData& ft) {
    ifCutTraj_ = ft.if_cut;
    unOccupied_traj_num_ = -1;
    TrajNum = (int)ft.UnOccupied_traj_pts.size() + 1;
    Innerpoints.assign(size_t(2) * (TrajNum - 1), 0.0);
    for (int i = 0; i < TrajNum - 1; i++) {
      Innerpoints[2 * i] = ft.UnOccupied_traj_pts[i][0];
      Innerpoints[2 * i + 1] = ft.UnOccupied_traj_pts[i][1];
    }
    inner_init_positions = ft.UnOccupied_positions;
    inner_init_positions.push_back({ft.final_state_XYTheta[0], ft.final_state_XYTheta[1], ft.final_state_XYTheta[2]});
    std::memcpy(iniState, ft.start_state, sizeof(iniState));
    std::memcpy(finState, ft.final_state, sizeof(finState));
    pieceTime.assign(TrajNum, 1.0);
    for (auto& t : pieceTime) t *= ft.UnOccupied_initT;
    std::memcpy(iniStateXYTheta, ft.start_state_XYTheta, sizeof(iniStateXYTheta));
    std::memcpy(finStateXYTheta, ft.final_state_XYTheta, sizeof(finStateXYTheta));
  }

  void unpackX(const Vec& x) {  // opt:346-357 / 452-464
    finalInnerpoints.assign(x.begin(), x.begin() + 2 * (TrajNum - 1));
    finState[1][0] = x[2 * (TrajNum - 1)];
    VirtualT2RealT(x.data() + 2 * (TrajNum - 1) + 1, TrajNum, finalpieceTime);
    Minco.setTConditions(finState);
    Minco.setParameters(finalInnerpoints.data(), finalpieceTime.data());
  }

  // opt:251-472.  skip_stage_a / skip_stage_b are test hooks (default: reference behaviour).
  bool optimizer(bool run_stage_a = true, bool run_stage_b = true) {
    if (!ifCutTraj_) {
      EqualLambda[0] = p.EqualLambda[0]; EqualLambda[1] = p.EqualLambda[1];
      EqualRho[0] = p.EqualRho[0]; EqualRho[1] = p.EqualRho[1];
    } else {
      EqualLambda[0] = p.CutEqualLambda[0]; EqualLambda[1] = p.CutEqualLambda[1];
      EqualRho[0] = p.CutEqualRho[0]; EqualRho[1] = p.CutEqualRho[1];
    }
    int variable_num_ = 3 * TrajNum - 1;
    Minco.setConditions(iniState, finState, TrajNum, p.energyWeights);
    Minco.setParameters(Innerpoints.data(), pieceTime.data());
    Vec x(variable_num_);
    int offset = 0;
    std::memcpy(x.data() + offset, Innerpoints.data(), Innerpoints.size() * sizeof(double));
    offset += (int)Innerpoints.size();
    x[offset] = finState[1][0];
    ++offset;
    RealT2VirtualT(pieceTime, x.data() + offset);
    double cost = 0.0;
    int result = 0;
    Vec g(x.size(), 0.0);
    iter_num_ = 0;
    alore_lbfgs_params_t path_params = p.path_lbfgs;
    if (std::fabs(finState[1][0]) < p.shot_path_horizon) path_params.past = p.shot_path_past;
    else path_params.past = p.normal_past;
    if (run_stage_a) {
      result = lbfgs_optimize(x, cost, [&](const Vec& xx, Vec& gg) { return costFunction(0, xx, gg); }, path_params);
      costFunction(0, x, g);  // opt:341 — the reference's extra (printing) evaluation mutates planner state
    }
    unpackX(x);
    iter_num_ = 0;
    last_alm_iters = 0;
    if (run_stage_b) {
      while (true) {
        result = lbfgs_optimize(x, cost, [&](const Vec& xx, Vec& gg) { return costFunction(1, xx, gg); }, p.lbfgs);
        last_alm_iters++;
        const double nrm = std::sqrt(FinalIntegralXYError[0] * FinalIntegralXYError[0] + FinalIntegralXYError[1] * FinalIntegralXYError[1]);
        if (!ifCutTraj_) {
          if (nrm < p.EqualTolerance[0]) break;
          EqualLambda[0] += EqualRho[0] * FinalIntegralXYError[0];
          EqualLambda[1] += EqualRho[1] * FinalIntegralXYError[1];
          EqualRho[0] = std::min((1 + p.EqualGamma[0]) * EqualRho[0], p.EqualRhoMax[0]);
          EqualRho[1] = std::min((1 + p.EqualGamma[1]) * EqualRho[1], p.EqualRhoMax[1]);
        } else {
          if (nrm < p.CutEqualTolerance[0]) break;
          EqualLambda[0] += EqualRho[0] * FinalIntegralXYError[0];
          EqualLambda[1] += EqualRho[1] * FinalIntegralXYError[1];
          EqualRho[0] = std::min((1 + p.CutEqualGamma[0]) * EqualRho[0], p.CutEqualRhoMax[0]);
          EqualRho[1] = std::min((1 + p.CutEqualGamma[1]) * EqualRho[1], p.CutEqualRhoMax[1]);
        }
        // The reference loops `while(ros::ok())` with no cap; alm_max_outer (0 = hard cap only)
        // is applied identically on the device.
        const int cap = p.alm_max_outer > 0 ? std::min(p.alm_max_outer, ALORE_ALM_HARD_CAP) : ALORE_ALM_HARD_CAP;
        if (last_alm_iters >= cap) break;
      }
    }
    unpackX(x);
    last_status = result;
    last_cost = cost;
    return true;
  }

  // opt:474-571
  bool check_final_collision(const Traj& final_traj, const double start_state_XYTheta[3], double* min_dist_out = nullptr) const {
    double ini_x = start_state_XYTheta[0], ini_y = start_state_XYTheta[1];
    double s1;
    int sparseResolution = p.finalSafeDisCheckNum;
    int SamNum = 2 * sparseResolution;
    double sumT = 0.0;
    int TrajN = final_traj.N;
    const Vec& pT = final_traj.T;
    std::vector<Vec> VecIntegralX(TrajN), VecIntegralY(TrajN);
    const double icr = p.ICR[2];
    for (int i = 0; i < TrajN; i++) {
      double step = pT[i] / sparseResolution;
      double halfstep = step / 2.0;
      double CoeffIntegral = pT[i] / sparseResolution / 6.0;
      Vec IntegralX(sparseResolution, 0.0), IntegralY(sparseResolution, 0.0);
      s1 = 0.0;
      for (int j = 0; j <= SamNum; j++) {
        double currPos[2], currVel[2];
        final_traj.getPos(s1 + sumT, currPos);
        final_traj.getVel(s1 + sumT, currVel);
        s1 += halfstep;
        if (p.if_standard_diff) {
          if (j % 2 == 0) {
            if (j != 0) {
              IntegralX[j / 2 - 1] += CoeffIntegral * currVel[1] * ocos(currPos[0]);
              IntegralY[j / 2 - 1] += CoeffIntegral * currVel[1] * osin(currPos[0]);
            }
            if (j != SamNum) {
              IntegralX[j / 2] += CoeffIntegral * currVel[1] * ocos(currPos[0]);
              IntegralY[j / 2] += CoeffIntegral * currVel[1] * osin(currPos[0]);
            }
          } else {
            IntegralX[j / 2] += 4.0 * CoeffIntegral * currVel[1] * ocos(currPos[0]);
            IntegralY[j / 2] += 4.0 * CoeffIntegral * currVel[1] * osin(currPos[0]);
          }
        } else {
          double cosyaw = ocos(currPos[0]), sinyaw = osin(currPos[0]);
          if (j % 2 == 0) {
            if (j != 0) {
              IntegralX[j / 2 - 1] += CoeffIntegral * (currVel[1] * cosyaw + currVel[0] * icr * sinyaw);
              IntegralY[j / 2 - 1] += CoeffIntegral * (currVel[1] * sinyaw - currVel[0] * icr * cosyaw);
            }
            if (j != SamNum) {
              IntegralX[j / 2] += CoeffIntegral * (currVel[1] * cosyaw + currVel[0] * icr * sinyaw);
              IntegralY[j / 2] += CoeffIntegral * (currVel[1] * sinyaw - currVel[0] * icr * cosyaw);
            }
          } else {
            IntegralX[j / 2] += 4.0 * CoeffIntegral * (currVel[1] * cosyaw + currVel[0] * icr * sinyaw);
            IntegralY[j / 2] += 4.0 * CoeffIntegral * (currVel[1] * sinyaw - currVel[0] * icr * cosyaw);
          }
        }
      }
      VecIntegralX[i] = IntegralX;
      VecIntegralY[i] = IntegralY;
      sumT += pT[i];
    }
    double min_distance = DBL_MAX;
    double pos[2] = {ini_x, ini_y};
    for (size_t i = 0; i < VecIntegralX.size(); i++)
      for (size_t j = 0; j < VecIntegralX[i].size(); j++) {
        pos[0] += VecIntegralX[i][j];
        pos[1] += VecIntegralY[i][j];
        double SDFvalue = map_->getDistWithGradBilinear(pos);
        if (SDFvalue < min_distance) min_distance = SDFvalue;
        if (SDFvalue < p.finalMinSafeDis) {
          if (min_dist_out) *min_dist_out = min_distance;
          return true;
        }
      }
    if (min_dist_out) *min_dist_out = min_distance;
    return false;
  }

  void getTrajectory(Traj& tr) const {  // minco:900-913 (coefficients kept in ascending order)
    tr.N = TrajNum;
    tr.T = Minco.T1;
    tr.coef = Minco.b;
  }

  // opt:169-220
  bool minco_plan(const FlatTrajData& flat_traj) {
    bool final_collision = false;
    int replan_num_for_coll = 0;
    double start_safe_dis = map_->getDistanceReal(flat_traj.start_state_XYTheta) * 0.85;
    safeDis = std::min(start_safe_dis, p.safeDis);
    total_evals = 0;
    for (; replan_num_for_coll < p.safeReplanMaxTime; replan_num_for_coll++) {
      get_state(flat_traj);
      optimizer();
      getTrajectory(optimizer_traj_);
      final_collision = check_final_collision(optimizer_traj_, iniStateXYTheta);
      if (final_collision) time_weight *= 0.75;
      else break;
    }
    time_weight = p.pw_time;
    safeDis = p.safeDis;
    replans = std::min(replan_num_for_coll + 1, p.safeReplanMaxTime);
    if (replan_num_for_coll == p.safeReplanMaxTime) return false;
    final_traj_ = optimizer_traj_;
    return true;
  }
};

// =====================================================================================
// Front-end time allocation that PRODUCES FlatTrajData (input generator for the candidates)
//                                                                        jps:217-441
// =====================================================================================
struct FrontEnd {
  double max_vel_ = 3.0, max_acc_ = 2.0;
  double yaw_weight_ = 0.30, distance_weight_ = 1.40;   // front_end/config/jps3ms.yaml
  double trajCutLength_ = 600.0, sampletime_ = 0.4;      // global_planning3ms.yaml: trajCutLength, timeResolution
  int mintrajNum_ = 3;

  static void normalizeAngle(double ref_angle, double& angle) {  // jps:368-375
    while (ref_angle - angle > M_PI) angle += 2 * M_PI;
    while (ref_angle - angle < -M_PI) angle -= 2 * M_PI;
  }
  double evaluateDuration(double length, double startV, double endV, double maxV, double maxA) const {  // jps:378-399
    double critical_len;
    double startv2 = std::pow(startV, 2), endv2 = std::pow(endV, 2), maxv2 = std::pow(maxV, 2);
    if (startV > maxV) startv2 = maxv2;
    if (endV > max_vel_) endv2 = maxv2;
    critical_len = (maxv2 - startv2) / (2 * maxA) + (maxv2 - endv2) / (2 * maxA);
    if (length >= critical_len) return (maxV - startV) / maxA + (maxV - endV) / maxA + (length - critical_len) / maxV;
    double tmpv = std::sqrt(0.5 * (startv2 + endv2 + 2 * maxA * length));
    return (tmpv - startV) / maxA + (tmpv - endV) / maxA;
  }
  double evaluateLength(double curt, double locallength, double /*localtime*/, double startV, double endV, double maxV, double maxA) const {  // jps:404-441
    double critical_len;
    double startv2 = std::pow(startV, 2), endv2 = std::pow(endV, 2), maxv2 = std::pow(maxV, 2);
    if (startV > maxV) startv2 = maxv2;
    if (endV > max_vel_) endv2 = maxv2;
    critical_len = (maxv2 - startv2) / (2 * maxA) + (maxv2 - endv2) / (2 * maxA);
    if (locallength >= critical_len) {
      double t1 = (maxV - startV) / maxA;
      double t2 = t1 + (locallength - critical_len) / maxV;
      if (curt <= t1) return startV * curt + 0.5 * maxA * std::pow(curt, 2);
      else if (curt <= t2) return startV * t1 + 0.5 * maxA * std::pow(t1, 2) + (curt - t1) * maxV;
      else return startV * t1 + 0.5 * maxA * std::pow(t1, 2) + (t2 - t1) * maxV + maxV * (curt - t2) - 0.5 * maxA * std::pow(curt - t2, 2);
    } else {
      double tmpv = std::sqrt(0.5 * (startv2 + endv2 + 2 * maxA * locallength));
      double tmpt = (tmpv - startV) / maxA;
      if (curt <= tmpt) return startV * curt + 0.5 * maxA * std::pow(curt, 2);
      else return startV * tmpt + 0.5 * maxA * std::pow(tmpt, 2) + tmpv * (curt - tmpt) - 0.5 * maxA * std::pow(curt - tmpt, 2);
    }
  }

  // path: >= 2 xy way-points (Unoccupied_path_); start/end state: x, y, yaw.
  // current_state_VAJ/OAJ: (v, a, j) and (omega, alpha, .) at the start (zeros for a parked robot).
  FlatTrajData make(const std::vector<std::array<double, 2>>& path, const double start_state_[3], const double end_state_[3],
                    const double VAJ[3], const double OAJ[3]) const {
    using S5 = std::array<double, 5>;  // x y theta dtheta ds
    std::vector<S5> samp;
    double cur_theta;
    samp.push_back({start_state_[0], start_state_[1], start_state_[2], 0, 0});                // jps:217-262
    cur_theta = std::atan2(path[1][1] - path[0][1], path[1][0] - path[0][0]);
    normalizeAngle(start_state_[2], cur_theta);
    samp.push_back({start_state_[0], start_state_[1], cur_theta, cur_theta - start_state_[2], 0});
    cur_theta = std::atan2(path[0][1] - path[1][1], path[0][0] - path[1][0]) + M_PI;
    normalizeAngle(start_state_[2], cur_theta);
    samp.push_back({start_state_[0], start_state_[1], cur_theta, cur_theta - start_state_[2], 0});
    int path_size = (int)path.size();
    for (int i = 1; i < path_size - 1; i++) {
      const auto& pt = path[i];
      const S5 bk = samp.back();
      samp.push_back({pt[0], pt[1], bk[2], 0, std::sqrt(std::pow(pt[0] - bk[0], 2) + std::pow(pt[1] - bk[1], 2))});
      cur_theta = std::atan2(path[i + 1][1] - path[i][1], path[i + 1][0] - path[i][0]);
      normalizeAngle(samp.back()[2], cur_theta);
      const S5 bk2 = samp.back();
      samp.push_back({pt[0], pt[1], cur_theta, cur_theta - bk2[2], 0});
    }
    {
      const auto& pt = path.back();
      const S5 bk = samp.back();
      samp.push_back({pt[0], pt[1], bk[2], 0, std::sqrt(std::pow(pt[0] - bk[0], 2) + std::pow(pt[1] - bk[1], 2))});
      cur_theta = end_state_[2];
      normalizeAngle(samp.back()[2], cur_theta);
      const S5 bk2 = samp.back();
      samp.push_back({pt[0], pt[1], cur_theta, cur_theta - bk2[2], 0});
    }
    // getTrajsWithTime                                                         jps:264-366
    std::vector<S5> cut;
    std::vector<double> pathlengths, wlengths;
    double AllW = 0, AllLen = 0;
    bool if_cut = false;
    double cut_state[3] = {samp.back()[0], samp.back()[1], samp.back()[2]};
    int PathNodeNum = (int)samp.size();
    cut.push_back(samp[0]);
    pathlengths.push_back(0);
    wlengths.push_back(0);
    for (int idx = 1; idx < PathNodeNum && !if_cut; idx++) {
      const S5& pn = samp[idx];
      if (AllLen + std::fabs(pn[4]) >= trajCutLength_ && pn[4] != 0) {
        if_cut = true;
        const S5& fs = samp[idx - 1];
        const double fr = (trajCutLength_ - AllLen) / std::fabs(pn[4]);
        for (int t = 0; t < 3; t++) cut_state[t] = fs[t] + (pn[t] - fs[t]) * (trajCutLength_ - AllLen) / std::fabs(pn[4]);
        S5 s5 = {cut_state[0], cut_state[1], cut_state[2], fr * pn[3], trajCutLength_ - AllLen};
        cut.push_back(s5);
        AllLen += s5[4];
        pathlengths.push_back(AllLen);
        AllW += yaw_weight_ * std::fabs(s5[3]) + distance_weight_ * std::fabs(s5[4]);
        wlengths.push_back(AllW);
        break;
      }
      cut.push_back(pn);
      AllLen += pn[4];
      pathlengths.push_back(AllLen);
      AllW += yaw_weight_ * std::fabs(pn[3]) + distance_weight_ * std::fabs(pn[4]);
      wlengths.push_back(AllW);
    }
    double totalT = evaluateDuration(AllW, VAJ[0], 0.0, max_vel_, max_acc_);
    FlatTrajData ft;
    double sampletime = totalT / std::max(int(totalT / sampletime_ + 0.5), mintrajNum_);
    int nodeIndex = 1;
    PathNodeNum = (int)cut.size();
    for (double samplet = sampletime; samplet < totalT - 1e-3; samplet += sampletime) {
      double arc = evaluateLength(samplet, AllW, totalT, VAJ[0], 0.0, max_vel_, max_acc_);
      for (int k = nodeIndex; k < PathNodeNum; k++) {
        const S5& pn = cut[k];
        const S5& pp = cut[k - 1];
        double tmparc = wlengths[k];
        if (tmparc >= arc) {
          nodeIndex = k;
          double l1 = tmparc - arc;
          double l = wlengths[k] - wlengths[k - 1];
          double interp_s = pathlengths[k - 1] + (l - l1) / l * (pn[4]);
          double interp_yaw = cut[k - 1][2] + (l - l1) / l * (pn[3]);
          ft.UnOccupied_traj_pts.push_back({interp_yaw, interp_s, samplet});
          double interp_x = l1 / l * pp[0] + (l - l1) / l * (pn[0]);
          double interp_y = l1 / l * pp[1] + (l - l1) / l * (pn[1]);
          ft.UnOccupied_positions.push_back({interp_x, interp_y, interp_yaw});
          break;
        }
      }
    }
    ft.start_state[0][0] = cut[0][2]; ft.start_state[1][0] = 0;
    ft.start_state[0][1] = OAJ[0]; ft.start_state[0][2] = OAJ[1];
    ft.start_state[1][1] = VAJ[0]; ft.start_state[1][2] = VAJ[1];
    ft.final_state[0][0] = cut[PathNodeNum - 1][2]; ft.final_state[1][0] = pathlengths[PathNodeNum - 1];
    ft.final_state[0][1] = ft.final_state[0][2] = ft.final_state[1][1] = ft.final_state[1][2] = 0.0;
    ft.UnOccupied_initT = sampletime;
    for (int t = 0; t < 3; t++) { ft.start_state_XYTheta[t] = start_state_[t]; ft.final_state_XYTheta[t] = cut_state[t]; }
    ft.if_cut = if_cut;
    return ft;
  }
};

}  // namespace orc
