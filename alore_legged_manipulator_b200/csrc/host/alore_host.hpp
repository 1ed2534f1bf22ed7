// alore_host.hpp — C++ host side above the C ABI: drop-in mirrors of the reference's grid-map and back_end
// classes for the hot path, same method names / argument meaning / error behaviour.
//
//   SDFmap     planning_ddr_opt/utils/plan_env/include/plan_env/sdf_map.h:96-262
//   MSPlanner  planning_ddr_opt/back_end/include/back_end/optimizer.h:192-300
//
// Eigen-free (the image has no Eigen): points are std::array<double,N>.  In the reference tree the same
// forwarding bodies go into the existing classes unchanged in signature — see INTEGRATION.md.
#pragma once
#include <algorithm>
#include <array>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/alore_b200.h"

namespace alore {

using Vec2d = std::array<double, 2>;
using Vec3d = std::array<double, 3>;
using Vec2i = std::array<int, 2>;

class Context {
 public:
  explicit Context(int device = 0) {
    if (alore_create(device, &h_) != ALORE_OK) throw std::runtime_error(std::string("alore_create: ") + alore_last_error(nullptr));
  }
  ~Context() { alore_destroy(h_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  alore_ctx* get() const { return h_; }
  void check(int rc) const { if (rc != ALORE_OK) throw std::runtime_error(alore_last_error(h_)); }
 private:
  alore_ctx* h_ = nullptr;
};

// FlatTrajData, planning_ddr_opt/front_end/include/front_end/traj_representation.h:46-76
struct FlatTrajData {
  std::vector<Vec3d> UnOccupied_traj_pts;   // yaw, s, t
  double UnOccupied_initT = 0.0;
  std::vector<Vec3d> UnOccupied_positions;  // x, y, yaw
  double start_state[2][3] = {{0}};         // rows yaw / s, columns P V A
  double final_state[2][3] = {{0}};
  Vec3d start_state_XYTheta{}, final_state_XYTheta{};
  bool if_cut = false;
};

class SDFmap {
 public:
  enum { Unknown, Unoccupied, Occupied };   // sdf_map.h:98
  double grid_interval_, inv_grid_interval_;
  double global_x_upper_, global_y_upper_, global_x_lower_, global_y_lower_;
  int GLX_SIZE_, GLY_SIZE_, GLXY_SIZE_;
  Vec3d odom_pos_{};
  bool ref_compat = true;

  // the ROS parameters of sdf_map.h:120-134 become constructor arguments
  SDFmap(Context& ctx, double gridmap_interval, double detection_range, double x_lower, double x_upper, double y_lower, double y_upper)
      : grid_interval_(gridmap_interval), inv_grid_interval_(1 / gridmap_interval), global_x_upper_(x_upper), global_y_upper_(y_upper),
        global_x_lower_(x_lower), global_y_lower_(y_lower), ctx_(ctx), detection_range_(detection_range) {
    GLX_SIZE_ = (int)std::ceil((global_x_upper_ - global_x_lower_) / grid_interval_);   // sdf_map.h:150-152
    GLY_SIZE_ = (int)std::ceil((global_y_upper_ - global_y_lower_) / grid_interval_);
    GLXY_SIZE_ = GLX_SIZE_ * GLY_SIZE_;
    gridmap_ = new uint8_t[GLXY_SIZE_];
    std::fill_n(gridmap_, GLXY_SIZE_, (uint8_t)Unknown);
    distance_buffer_all_.assign(GLXY_SIZE_, std::numeric_limits<double>::max());
    const alore_map_geom_t g = geom();
    ctx_.check(alore_esdf_reset(ctx_.get(), &g));
    // the map owns both buffers for its whole life -> page-lock them once (optional, best effort)
    pinned_occ_ = alore_host_register(ctx_.get(), gridmap_, (size_t)GLXY_SIZE_) == ALORE_OK;
    pinned_dist_ = alore_host_register(ctx_.get(), distance_buffer_all_.data(), sizeof(double) * (size_t)GLXY_SIZE_) == ALORE_OK;
  }
  ~SDFmap() {
    if (pinned_occ_) alore_host_unregister(ctx_.get(), gridmap_);
    if (pinned_dist_) alore_host_unregister(ctx_.get(), distance_buffer_all_.data());
    delete[] gridmap_;
  }
  SDFmap(const SDFmap&) = delete;
  SDFmap& operator=(const SDFmap&) = delete;

  alore_map_geom_t geom() const {
    return alore_map_geom_t{GLX_SIZE_, GLY_SIZE_, global_x_lower_, global_y_lower_, global_x_upper_, global_y_upper_, grid_interval_, inv_grid_interval_};
  }
  uint8_t* gridmap() { return gridmap_; }
  std::vector<double>& distance_buffer_all() { return distance_buffer_all_; }
  int Index2Vectornum(int x, int y) const { return x * GLY_SIZE_ + y; }                 // sdf_map.cpp:525-527
  Vec2d gridIndex2coordd(int x, int y) const {                                          // sdf_map.cpp:460-465
    return {((double)x + 0.5) * grid_interval_ + global_x_lower_, ((double)y + 0.5) * grid_interval_ + global_y_lower_};
  }
  Vec2i coord2gridIndex(const Vec2d& pt) const {                                        // sdf_map.cpp:467-472
    return {std::min(std::max(int((pt[0] - global_x_lower_) * inv_grid_interval_), 0), GLX_SIZE_ - 1),
            std::min(std::max(int((pt[1] - global_y_lower_) * inv_grid_interval_), 0), GLY_SIZE_ - 1)};
  }
  void setObs(const Vec2d& coord) { paint(coord, Occupied); }                           // sdf_map.cpp:485-496
  void setFree(const Vec2d& coord) { paint(coord, Unoccupied); }                        // sdf_map.cpp:498-509
  bool isOccupied(int ix, int iy) const { return gridmap_[Index2Vectornum(ix, iy)] == Occupied; }

  // sdf_map.cpp:618-680: the window is computed with the reference's own expression, the body runs on the GPU
  void updateESDF2d() {
    const int min_x = (int)std::floor(std::max(0.0, odom_pos_[0] - detection_range_ - global_x_lower_) * inv_grid_interval_);
    const int min_y = (int)std::floor(std::max(0.0, odom_pos_[1] - detection_range_ - global_y_lower_) * inv_grid_interval_);
    const int max_x = (int)(std::ceil(std::min(global_x_upper_ - global_x_lower_, odom_pos_[0] + detection_range_ - global_x_lower_) * inv_grid_interval_) - 1);
    const int max_y = (int)(std::ceil(std::min(global_y_upper_ - global_y_lower_, odom_pos_[1] + detection_range_ - global_y_lower_) * inv_grid_interval_) - 1);
    const alore_map_geom_t g = geom();
    ctx_.check(alore_esdf_update(ctx_.get(), &g, gridmap_, min_x, min_y, max_x, max_y, distance_buffer_all_.data(), ref_compat ? 1 : 0));
  }
  void forceUpdateESDF() {                                                              // sdf_map.cpp:511-516
    if (!has_map_) return;
    esdf_need_update_ = true;
    updateESDF2d();
    has_esdf_ = true;
  }
  double getDistanceReal(const Vec2d& pos) const {                                      // sdf_map.cpp:865-871
    if (out_of_map(pos)) return 10000;
    const Vec2i idx = coord2gridIndex(pos);
    return distance_buffer_all_[(size_t)idx[0] * GLY_SIZE_ + idx[1]];
  }
  bool isOccWithSafeDis(int ix, int iy, double safe_dis) const {                        // sdf_map.cpp:946-948
    return distance_buffer_all_[(size_t)Index2Vectornum(ix, iy)] < safe_dis;
  }
  double getDistWithGradBilinear(const Vec2d& pos, Vec2d& grad) const {                 // sdf_map.cpp:760-794 (host mirror)
    if (out_of_map(pos)) { grad = {0, 0}; return 100; }
    int ix = std::min(std::max(int((pos[0] - global_x_lower_) * inv_grid_interval_ - 0.5), 0), GLX_SIZE_ - 1);
    int iy = std::min(std::max(int((pos[1] - global_y_lower_) * inv_grid_interval_ - 0.5), 0), GLY_SIZE_ - 1);
    if (ix >= GLX_SIZE_ - 1 || iy >= GLY_SIZE_ - 1) { grad = {0, 0}; return 100; }
    const Vec2d c = gridIndex2coordd(ix, iy);
    const double dx = (pos[0] - c[0]) * inv_grid_interval_, dy = (pos[1] - c[1]) * inv_grid_interval_;
    const double* d = distance_buffer_all_.data() + (size_t)ix * GLY_SIZE_ + iy;
    const double v00 = d[0], v01 = d[1], v10 = d[GLY_SIZE_], v11 = d[GLY_SIZE_ + 1];
    const double v0 = (1 - dx) * v00 + dx * v10, v1 = (1 - dx) * v01 + dx * v11;
    grad[1] = (v1 - v0) * inv_grid_interval_;
    grad[0] = ((1 - dy) * (v10 - v00) + dy * (v11 - v01)) * inv_grid_interval_;
    return (1 - dy) * v0 + dy * v1;
  }
  bool has_map_ = false, has_esdf_ = false, esdf_need_update_ = false;

 private:
  bool out_of_map(const Vec2d& p) const { return p[0] < global_x_lower_ || p[1] < global_y_lower_ || p[0] > global_x_upper_ || p[1] > global_y_upper_; }
  void paint(const Vec2d& coord, uint8_t state) {
    const float cx = (float)coord[0], cy = (float)coord[1];   // the reference truncates through float
    if (cx < global_x_lower_ || cy < global_y_lower_ || cx >= global_x_upper_ || cy >= global_y_upper_) return;
    const int ix = static_cast<int>((cx - global_x_lower_) * inv_grid_interval_);
    const int iy = static_cast<int>((cy - global_y_lower_) * inv_grid_interval_);
    gridmap_[ix * GLY_SIZE_ + iy] = state;
    has_map_ = true;
    esdf_need_update_ = true;
  }
  Context& ctx_;
  double detection_range_;
  uint8_t* gridmap_ = nullptr;                   // sdf_map.h:73
  std::vector<double> distance_buffer_all_;     // sdf_map.h:69
  bool pinned_occ_ = false, pinned_dist_ = false;
};

// Result of one candidate (what MSPlanner exposes through final_traj_ and its getters).
struct PlanResult {
  bool ok = false;
  int status = 0, replans = 0, alm_iters = 0, evals = 0;
  double cost = 0.0, tail_s = 0.0;
  std::vector<double> inner_pts;   // 2 x (N-1) column-major (finalInnerpoints)
  std::vector<double> piece_T;     // finalpieceTime
  std::vector<double> coeffs;      // 6N x 2, ascending powers
};

class MSPlanner {
 public:
  alore_params_t params;
  MSPlanner(Context& ctx, SDFmap& map) : ctx_(ctx), map_(map) { alore_params_default(&params); }

  // optimizer.h:207 — same name, argument and bool result; results in final_result()
  bool minco_plan(const FlatTrajData& flat_traj) {
    std::vector<PlanResult> r = minco_plan_batch({flat_traj});
    last_ = r[0];
    return last_.ok;
  }
  const PlanResult& final_result() const { return last_; }

  // the batched call the task-and-motion planner uses: one FlatTrajData per candidate leg
  std::vector<PlanResult> minco_plan_batch(const std::vector<FlatTrajData>& fts) {
    const int B = (int)fts.size();
    std::vector<int32_t> po(B + 1, 0);
    for (int b = 0; b < B; b++) po[b + 1] = po[b] + (int)fts[b].UnOccupied_traj_pts.size() + 1;
    const int tot = po[B];
    std::vector<double> ip(2 * (size_t)std::max(tot - B, 1)), T(B), pos(3 * (size_t)tot), ss(6 * (size_t)B), fs(6 * (size_t)B), sx(3 * (size_t)B), fx(3 * (size_t)B);
    std::vector<uint8_t> cut(B);
    for (int b = 0; b < B; b++) {
      const FlatTrajData& f = fts[b];
      const int N = po[b + 1] - po[b];
      for (int i = 0; i < N - 1; i++) {
        ip[2 * (size_t)(po[b] - b + i)] = f.UnOccupied_traj_pts[i][0];
        ip[2 * (size_t)(po[b] - b + i) + 1] = f.UnOccupied_traj_pts[i][1];
        for (int t = 0; t < 3; t++) pos[3 * (size_t)(po[b] + i) + t] = f.UnOccupied_positions[i][t];
      }
      for (int t = 0; t < 3; t++) pos[3 * (size_t)(po[b] + N - 1) + t] = f.final_state_XYTheta[t];   // optimizer.cpp:234-235
      T[b] = f.UnOccupied_initT;
      for (int d = 0; d < 2; d++)
        for (int k = 0; k < 3; k++) { ss[6 * (size_t)b + 3 * d + k] = f.start_state[d][k]; fs[6 * (size_t)b + 3 * d + k] = f.final_state[d][k]; }
      for (int t = 0; t < 3; t++) { sx[3 * (size_t)b + t] = f.start_state_XYTheta[t]; fx[3 * (size_t)b + t] = f.final_state_XYTheta[t]; }
      cut[b] = f.if_cut ? 1 : 0;
    }
    alore_candidates_t c{B, po.data(), ip.data(), T.data(), pos.data(), ss.data(), fs.data(), sx.data(), fx.data(), cut.data()};
    std::vector<int32_t> ok(B), st(B), rp(B), al(B), ev(B);
    std::vector<double> cost(B), oip(2 * (size_t)std::max(tot - B, 1)), ts(B), oT(tot), oc(12 * (size_t)tot);
    alore_results_t r{ok.data(), st.data(), rp.data(), al.data(), ev.data(), cost.data(), oip.data(), ts.data(), oT.data(), oc.data()};
    ctx_.check(alore_opt_batch(ctx_.get(), &params, &c, &r));
    std::vector<PlanResult> out(B);
    for (int b = 0; b < B; b++) {
      const int N = po[b + 1] - po[b];
      PlanResult& p = out[b];
      p.ok = ok[b] != 0; p.status = st[b]; p.replans = rp[b]; p.alm_iters = al[b]; p.evals = ev[b]; p.cost = cost[b]; p.tail_s = ts[b];
      p.inner_pts.assign(oip.begin() + 2 * (size_t)(po[b] - b), oip.begin() + 2 * (size_t)(po[b] - b + N - 1));
      p.piece_T.assign(oT.begin() + po[b], oT.begin() + po[b + 1]);
      p.coeffs.assign(oc.begin() + 12 * (size_t)po[b], oc.begin() + 12 * (size_t)po[b + 1]);
    }
    return out;
  }

 private:
  Context& ctx_;
  SDFmap& map_;
  PlanResult last_;
};

}  // namespace alore
