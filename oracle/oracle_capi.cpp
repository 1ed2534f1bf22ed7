// oracle_capi.cpp — extern "C" surface of the CPU ORACLE for ctypes (tests/, smoke(), bench.py's
// cpu_baseline and --impl reference legs ONLY).  Test infrastructure; never linked by the product.
#include <atomic>
#include <chrono>
#include <thread>

#include "alore_oracle.hpp"

using namespace orc;

namespace {
FlatTrajData make_ft(const alore_candidates_t* c, int b) {
  FlatTrajData ft;
  const int p0 = c->piece_off[b], N = c->piece_off[b + 1] - p0;
  const double* ip = c->inner_pts + 2 * size_t(p0 - b);
  const double* pos = c->inner_init_pos + 3 * size_t(p0);
  for (int i = 0; i < N - 1; i++) {
    ft.UnOccupied_traj_pts.push_back({ip[2 * i], ip[2 * i + 1], 0.0});
    ft.UnOccupied_positions.push_back({pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]});
  }
  ft.UnOccupied_initT = c->init_T[b];
  std::memcpy(ft.start_state, c->start_state + 6 * size_t(b), sizeof(ft.start_state));
  std::memcpy(ft.final_state, c->final_state + 6 * size_t(b), sizeof(ft.final_state));
  std::memcpy(ft.start_state_XYTheta, c->start_xytheta + 3 * size_t(b), sizeof(ft.start_state_XYTheta));
  std::memcpy(ft.final_state_XYTheta, c->final_xytheta + 3 * size_t(b), sizeof(ft.final_state_XYTheta));
  ft.if_cut = c->if_cut[b] != 0;
  return ft;
}
SdfMap view_map(const alore_map_geom_t* geom, const double* dist) {
  SdfMap m;
  m.g = *geom;
  m.dist_view_ = dist;
  return m;
}
}  // namespace

extern "C" {

void orc_esdf_window(const alore_map_geom_t* geom, double odom_x, double odom_y, double range, int* mn, int* mx) {
  SdfMap m;
  m.g = *geom;
  m.window(odom_x, odom_y, range, mn, mx);
}

// Literal updateESDF2d on caller buffers.  sq_pos/sq_neg (may be NULL): pre-sqrt values at the
// reference's own buffer index x*update_Y_SIZE + y, arrays of (X+1)*(Y+1) doubles.
int orc_esdf_update(const alore_map_geom_t* geom, const uint8_t* occ, int min_x, int min_y, int max_x, int max_y,
                    double* dist_inout, double* sq_pos, double* sq_neg) {
  SdfMap m;
  m.g = *geom;
  int mn[2] = {min_x, min_y}, mx[2] = {max_x, max_y};
  std::vector<double> sp, sn;
  m.updateESDF2d(occ, dist_inout, mn, mx, sq_pos ? &sp : nullptr, sq_neg ? &sn : nullptr);
  if (sq_pos) std::memcpy(sq_pos, sp.data(), sp.size() * sizeof(double));
  if (sq_neg) std::memcpy(sq_neg, sn.data(), sn.size() * sizeof(double));
  return 0;
}

// Row-parallel variant for the CPU-all baseline: same arithmetic per row/column, rows and
// columns distributed over threads (the aliasing writes are reproduced by running the two
// conflicting indices in reference order afterwards).  Used for timing only.
double orc_esdf_update_timed(const alore_map_geom_t* geom, const uint8_t* occ, int min_x, int min_y, int max_x, int max_y,
                             double* dist_inout, int reps) {
  SdfMap m;
  m.g = *geom;
  int mn[2] = {min_x, min_y}, mx[2] = {max_x, max_y};
  double best = 1e300;
  for (int r = 0; r < reps; r++) {
    auto t0 = std::chrono::steady_clock::now();
    m.updateESDF2d(occ, dist_inout, mn, mx, nullptr, nullptr);
    auto t1 = std::chrono::steady_clock::now();
    best = std::min(best, std::chrono::duration<double>(t1 - t0).count());
  }
  return best;
}

double orc_dist_grad3(const alore_map_geom_t* geom, const double* dist, const double* pos, double* grad, double mindis) {
  SdfMap m = view_map(geom, dist);
  return m.getDistWithGradBilinear(pos, grad, mindis);
}
double orc_dist_grad2(const alore_map_geom_t* geom, const double* dist, const double* pos, double* grad) {
  SdfMap m = view_map(geom, dist);
  return m.getDistWithGradBilinear(pos, grad);
}
double orc_dist1(const alore_map_geom_t* geom, const double* dist, const double* pos) {
  SdfMap m = view_map(geom, dist);
  return m.getDistWithGradBilinear(pos);
}
double orc_dist_real(const alore_map_geom_t* geom, const double* dist, const double* pos) {
  SdfMap m = view_map(geom, dist);
  return m.getDistanceReal(pos);
}

// MINCO forward: head/tail [2][3], inPs 2x(N-1) col-major, T[N] -> coeffs [6N][2]; also energy & partials.
int orc_minco_solve(int N, const double* head, const double* tail, const double* inPs, const double* T, const double* ew,
                    double* coeffs, double* energy, double* gdC, double* gdT) {
  MincoS3NU mc;
  double h[2][3], t[2][3];
  std::memcpy(h, head, sizeof(h));
  std::memcpy(t, tail, sizeof(t));
  mc.setConditions(h, t, N, ew);
  mc.setParameters(inPs, T);
  std::memcpy(coeffs, mc.b.data(), mc.b.size() * sizeof(double));
  if (energy) mc.getEnergy(*energy);
  if (gdC) { std::vector<double> v; mc.getEnergyPartialGradByCoeffs(v); std::memcpy(gdC, v.data(), v.size() * sizeof(double)); }
  if (gdT) { std::vector<double> v; mc.getEnergyPartialGradByTimes(v); std::memcpy(gdT, v.data(), v.size() * sizeof(double)); }
  return 0;
}
// MINCO adjoint: partial gradC [6N][2], partial gradT [N] -> gradP 2x(N-1), gradT [N], gradTail [2].
int orc_minco_adjoint(int N, const double* head, const double* tail, const double* inPs, const double* T, const double* ew,
                      const double* pgC, const double* pgT, double* gradP, double* gradT, double* gradTail) {
  MincoS3NU mc;
  double h[2][3], t[2][3];
  std::memcpy(h, head, sizeof(h));
  std::memcpy(t, tail, sizeof(t));
  mc.setConditions(h, t, N, ew);
  mc.setParameters(inPs, T);
  std::vector<double> c(pgC, pgC + 12 * size_t(N)), tt(pgT, pgT + N), gp, gt;
  mc.propogateArcYawLenghGrad(c, tt, gp, gt, gradTail);
  std::memcpy(gradP, gp.data(), gp.size() * sizeof(double));
  std::memcpy(gradT, gt.data(), gt.size() * sizeof(double));
  return 0;
}
// Dense banded matrix A (6N x 6N, row-major) as setParameters builds it, for invariants tests.
int orc_minco_matrix(int N, const double* T, double* Adense) {
  const int n = 6 * N;
  std::fill(Adense, Adense + size_t(n) * n, 0.0);
  auto Aref = [&](int i, int j) -> double& { return Adense[size_t(i) * n + j]; };
  std::vector<double> T1(N), T2(N), T3(N), T4(N), T5(N);
  for (int i = 0; i < N; i++) { T1[i] = T[i]; T2[i] = T1[i] * T1[i]; T3[i] = T2[i] * T1[i]; T4[i] = T2[i] * T2[i]; T5[i] = T4[i] * T1[i]; }
  Aref(0, 0) = 1.0; Aref(1, 1) = 1.0; Aref(2, 2) = 2.0;
  for (int i = 0; i < N - 1; i++) {
    Aref(6 * i + 3, 6 * i + 3) = 6.0; Aref(6 * i + 3, 6 * i + 4) = 24.0 * T1[i]; Aref(6 * i + 3, 6 * i + 5) = 60.0 * T2[i]; Aref(6 * i + 3, 6 * i + 9) = -6.0;
    Aref(6 * i + 4, 6 * i + 4) = 24.0; Aref(6 * i + 4, 6 * i + 5) = 120.0 * T1[i]; Aref(6 * i + 4, 6 * i + 10) = -24.0;
    Aref(6 * i + 5, 6 * i) = 1.0; Aref(6 * i + 5, 6 * i + 1) = T1[i]; Aref(6 * i + 5, 6 * i + 2) = T2[i]; Aref(6 * i + 5, 6 * i + 3) = T3[i]; Aref(6 * i + 5, 6 * i + 4) = T4[i]; Aref(6 * i + 5, 6 * i + 5) = T5[i];
    Aref(6 * i + 6, 6 * i) = 1.0; Aref(6 * i + 6, 6 * i + 1) = T1[i]; Aref(6 * i + 6, 6 * i + 2) = T2[i]; Aref(6 * i + 6, 6 * i + 3) = T3[i]; Aref(6 * i + 6, 6 * i + 4) = T4[i]; Aref(6 * i + 6, 6 * i + 5) = T5[i]; Aref(6 * i + 6, 6 * i + 6) = -1.0;
    Aref(6 * i + 7, 6 * i + 1) = 1.0; Aref(6 * i + 7, 6 * i + 2) = 2 * T1[i]; Aref(6 * i + 7, 6 * i + 3) = 3 * T2[i]; Aref(6 * i + 7, 6 * i + 4) = 4 * T3[i]; Aref(6 * i + 7, 6 * i + 5) = 5 * T4[i]; Aref(6 * i + 7, 6 * i + 7) = -1.0;
    Aref(6 * i + 8, 6 * i + 2) = 2.0; Aref(6 * i + 8, 6 * i + 3) = 6 * T1[i]; Aref(6 * i + 8, 6 * i + 4) = 12 * T2[i]; Aref(6 * i + 8, 6 * i + 5) = 20 * T3[i]; Aref(6 * i + 8, 6 * i + 8) = -2.0;
  }
  Aref(6 * N - 3, 6 * N - 6) = 1.0; Aref(6 * N - 3, 6 * N - 5) = T1[N - 1]; Aref(6 * N - 3, 6 * N - 4) = T2[N - 1]; Aref(6 * N - 3, 6 * N - 3) = T3[N - 1]; Aref(6 * N - 3, 6 * N - 2) = T4[N - 1]; Aref(6 * N - 3, 6 * N - 1) = T5[N - 1];
  Aref(6 * N - 2, 6 * N - 5) = 1.0; Aref(6 * N - 2, 6 * N - 4) = 2 * T1[N - 1]; Aref(6 * N - 2, 6 * N - 3) = 3 * T2[N - 1]; Aref(6 * N - 2, 6 * N - 2) = 4 * T3[N - 1]; Aref(6 * N - 2, 6 * N - 1) = 5 * T4[N - 1];
  Aref(6 * N - 1, 6 * N - 4) = 2; Aref(6 * N - 1, 6 * N - 3) = 6 * T1[N - 1]; Aref(6 * N - 1, 6 * N - 2) = 12 * T2[N - 1]; Aref(6 * N - 1, 6 * N - 1) = 20 * T3[N - 1];
  return 0;
}

// attachPenaltyFunctional on given coefficients, from cost=0 / zero partial gradients
// (same contract as alore_penalty_batch), one trajectory.
int orc_penalty(const alore_params_t* prm, const alore_map_geom_t* geom, const double* dist, int N, const double* coeffs,
                const double* T, const double* start_xy, const double* final_xy, double* cost, double* gradC, double* gradT,
                double* xy_err) {
  SdfMap m = view_map(geom, dist);
  MSPlanner pl;
  pl.init(*prm, &m);
  pl.TrajNum = N;
  pl.pieceTime.assign(T, T + N);
  pl.Minco.N = N;
  pl.Minco.b.assign(coeffs, coeffs + 12 * size_t(N));
  pl.iniStateXYTheta[0] = start_xy[0]; pl.iniStateXYTheta[1] = start_xy[1];
  pl.finStateXYTheta[0] = final_xy[0]; pl.finStateXYTheta[1] = final_xy[1];
  pl.EqualLambda[0] = prm->EqualLambda[0]; pl.EqualLambda[1] = prm->EqualLambda[1];
  pl.EqualRho[0] = prm->EqualRho[0]; pl.EqualRho[1] = prm->EqualRho[1];
  pl.safeDis = prm->safeDis;
  pl.partialGradByCoeffs.assign(12 * size_t(N), 0.0);
  pl.partialGradByTimes.assign(N, 0.0);
  double c = 0.0;
  pl.attachPenaltyFunctional(c);
  *cost = c;
  std::memcpy(gradC, pl.partialGradByCoeffs.data(), 12 * size_t(N) * sizeof(double));
  std::memcpy(gradT, pl.partialGradByTimes.data(), size_t(N) * sizeof(double));
  xy_err[0] = pl.FinalIntegralXYError[0]; xy_err[1] = pl.FinalIntegralXYError[1];
  return 0;
}
int orc_penalty_batch(const alore_params_t* prm, const alore_map_geom_t* geom, const double* dist, int B, const int32_t* piece_off,
                      const double* coeffs, const double* T, const double* start_xy, const double* final_xy, double* cost,
                      double* gradC, double* gradT, double* xy_err, int nthreads) {
  std::atomic<int> next{0};
  auto work = [&]() {
    for (;;) {
      int b = next.fetch_add(1);
      if (b >= B) break;
      const int p0 = piece_off[b], N = piece_off[b + 1] - p0;
      orc_penalty(prm, geom, dist, N, coeffs + 12 * size_t(p0), T + p0, start_xy + 2 * size_t(b), final_xy + 2 * size_t(b),
                  cost + b, gradC + 12 * size_t(p0), gradT + p0, xy_err + 2 * size_t(b));
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < std::max(1, nthreads); t++) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
  return 0;
}

// One costFunctionCallback (stage 1) / costFunctionCallbackPath (stage 0) evaluation for
// candidate b of a batch, entered the way optimizer() enters it (opt:251-273).
int orc_cost(const alore_params_t* prm, const alore_map_geom_t* geom, const double* dist, const alore_candidates_t* cands,
             int b, int stage, const double* x, const double* lambda, const double* rho, double safe_dis, double* cost,
             double* g, double* xy_err) {
  SdfMap m = view_map(geom, dist);
  MSPlanner pl;
  pl.init(*prm, &m);
  FlatTrajData ft = make_ft(cands, b);
  pl.get_state(ft);
  pl.safeDis = safe_dis;
  if (lambda) { pl.EqualLambda[0] = lambda[0]; pl.EqualLambda[1] = lambda[1]; }
  else { pl.EqualLambda[0] = prm->EqualLambda[0]; pl.EqualLambda[1] = prm->EqualLambda[1]; }
  if (rho) { pl.EqualRho[0] = rho[0]; pl.EqualRho[1] = rho[1]; }
  else { pl.EqualRho[0] = prm->EqualRho[0]; pl.EqualRho[1] = prm->EqualRho[1]; }
  pl.Minco.setConditions(pl.iniState, pl.finState, pl.TrajNum, prm->energyWeights);
  const int n = 3 * pl.TrajNum - 1;
  Vec xx(x, x + n), gg(g, g + n);
  *cost = pl.costFunction(stage, xx, gg);
  std::memcpy(g, gg.data(), n * sizeof(double));
  if (xy_err) { xy_err[0] = pl.FinalIntegralXYError[0]; xy_err[1] = pl.FinalIntegralXYError[1]; }
  return 0;
}

// Initial decision vector of optimizer() (opt:277-286) for candidate b.
int orc_initial_x(const alore_candidates_t* cands, int b, double* x) {
  FlatTrajData ft = make_ft(cands, b);
  MSPlanner pl;
  pl.get_state(ft);
  std::memcpy(x, pl.Innerpoints.data(), pl.Innerpoints.size() * sizeof(double));
  size_t off = pl.Innerpoints.size();
  x[off++] = pl.finState[1][0];
  MSPlanner::RealT2VirtualT(pl.pieceTime, x + off);
  return 0;
}

// B x minco_plan, candidates distributed over nthreads host threads (CPU-all baseline).
// stages: bit0 = run stage A, bit1 = run stage B (3 = reference behaviour).
int orc_opt_batch(const alore_params_t* prm, const alore_map_geom_t* geom, const double* dist, const alore_candidates_t* cands,
                  alore_results_t* out, int nthreads) {
  SdfMap m = view_map(geom, dist);
  std::atomic<int> next{0};
  const int B = cands->B;
  auto work = [&]() {
    MSPlanner pl;
    pl.init(*prm, &m);
    for (;;) {
      int b = next.fetch_add(1);
      if (b >= B) break;
      FlatTrajData ft = make_ft(cands, b);
      bool ok = pl.minco_plan(ft);
      const int p0 = cands->piece_off[b], N = cands->piece_off[b + 1] - p0;
      if (out->ok) out->ok[b] = ok ? 1 : 0;
      if (out->status) out->status[b] = pl.last_status;
      if (out->replans) out->replans[b] = pl.replans;
      if (out->alm_iters) out->alm_iters[b] = pl.last_alm_iters;
      if (out->evals) out->evals[b] = (int)pl.total_evals;
      if (out->cost) out->cost[b] = pl.last_cost;
      if (out->inner_pts) std::memcpy(out->inner_pts + 2 * size_t(p0 - b), pl.finalInnerpoints.data(), 2 * size_t(N - 1) * sizeof(double));
      if (out->tail_s) out->tail_s[b] = pl.finState[1][0];
      if (out->piece_T) std::memcpy(out->piece_T + p0, pl.finalpieceTime.data(), size_t(N) * sizeof(double));
      if (out->coeffs) std::memcpy(out->coeffs + 12 * size_t(p0), pl.Minco.b.data(), 12 * size_t(N) * sizeof(double));
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < std::max(1, nthreads); t++) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
  return 0;
}

int orc_final_collision(const alore_params_t* prm, const alore_map_geom_t* geom, const double* dist, int N, const double* coeffs,
                        const double* T, const double* start_xy, int32_t* collided, double* min_dist) {
  SdfMap m = view_map(geom, dist);
  MSPlanner pl;
  pl.init(*prm, &m);
  Traj tr;
  tr.N = N;
  tr.T.assign(T, T + N);
  tr.coef.assign(coeffs, coeffs + 12 * size_t(N));
  double st[3] = {start_xy[0], start_xy[1], 0.0};
  double md = 0.0;
  *collided = pl.check_final_collision(tr, st, &md) ? 1 : 0;
  if (min_dist) *min_dist = md;
  return 0;
}

// L-BFGS known-answer hook: minimise the extended Rosenbrock function from x (in/out).
int orc_lbfgs_rosenbrock(int n, double* x, const alore_lbfgs_params_t* param, double* f, int* iters, int* evals) {
  Vec xx(x, x + n);
  int ev = 0;
  auto fn = [&](const Vec& v, Vec& g) {
    ev++;
    double fx = 0.0;
    for (int i = 0; i < n; i += 2) {
      double t1 = 1.0 - v[i];
      double t2 = 10.0 * (v[i + 1] - v[i] * v[i]);
      g[i + 1] = 20.0 * t2;
      g[i] = -2.0 * (v[i] * g[i + 1] + t1);
      fx += t1 * t1 + t2 * t2;
    }
    return fx;
  };
  int it = 0;
  int ret = lbfgs_optimize(xx, *f, fn, *param, &it);
  std::memcpy(x, xx.data(), n * sizeof(double));
  if (iters) *iters = it;
  if (evals) *evals = ev;
  return ret;
}

// Front-end time allocation (jps:217-366): way-points -> one candidate.  Returns N (pieces),
// or -1 if capacity (max_pieces) is too small.  Output arrays follow alore_candidates_t rows.
int orc_frontend_make(int n_path, const double* path_xy, const double* start_xyt, const double* end_xyt, const double* VAJ,
                      const double* OAJ, double max_vel, double max_acc, double yaw_weight, double distance_weight,
                      double traj_cut_length, double sample_time, int min_traj_num, int max_pieces, double* inner_pts,
                      double* init_T, double* inner_init_pos, double* start_state, double* final_state, double* final_xytheta,
                      uint8_t* if_cut) {
  FrontEnd fe;
  fe.max_vel_ = max_vel; fe.max_acc_ = max_acc; fe.yaw_weight_ = yaw_weight; fe.distance_weight_ = distance_weight;
  fe.trajCutLength_ = traj_cut_length; fe.sampletime_ = sample_time; fe.mintrajNum_ = min_traj_num;
  std::vector<std::array<double, 2>> path(n_path);
  for (int i = 0; i < n_path; i++) path[i] = {path_xy[2 * i], path_xy[2 * i + 1]};
  FlatTrajData ft = fe.make(path, start_xyt, end_xyt, VAJ, OAJ);
  const int N = (int)ft.UnOccupied_traj_pts.size() + 1;
  if (N > max_pieces) return -1;
  for (int i = 0; i < N - 1; i++) {
    inner_pts[2 * i] = ft.UnOccupied_traj_pts[i][0];
    inner_pts[2 * i + 1] = ft.UnOccupied_traj_pts[i][1];
    for (int t = 0; t < 3; t++) inner_init_pos[3 * i + t] = ft.UnOccupied_positions[i][t];
  }
  for (int t = 0; t < 3; t++) inner_init_pos[3 * (N - 1) + t] = ft.final_state_XYTheta[t];
  *init_T = ft.UnOccupied_initT;
  std::memcpy(start_state, ft.start_state, sizeof(ft.start_state));
  std::memcpy(final_state, ft.final_state, sizeof(ft.final_state));
  for (int t = 0; t < 3; t++) final_xytheta[t] = ft.final_state_XYTheta[t];
  *if_cut = ft.if_cut ? 1 : 0;
  return N;
}

// The three analytic problems of oracle/ref_lbfgs_capi.cpp through the ORACLE's lbfgs_optimize (same callbacks).
int orc_lbfgs_run(int kind, int n, double* x, const alore_lbfgs_params_t* param, double* f, int* evals) {
  Vec xx(x, x + n);
  int ev = 0;
  auto fn = [&](const Vec& v, Vec& g) {
    ev++;
    double fx = 0.0;
    if (kind == 0) {
      for (int i = 0; i < n; i += 2) {
        double t1 = 1.0 - v[i];
        double t2 = 10.0 * (v[i + 1] - v[i] * v[i]);
        g[i + 1] = 20.0 * t2;
        g[i] = -2.0 * (v[i] * g[i + 1] + t1);
        fx += t1 * t1 + t2 * t2;
      }
    } else {
      for (int i = 0; i < n; i++) {
        const double w = 1.0 + 50.0 * i;
        const double a = v[i] - 0.1 * i;
        fx += 0.5 * w * a * a + (kind == 1 ? std::fabs(v[i]) : 0.0);
        g[i] = w * a + (kind == 1 ? (v[i] > 0 ? 1.0 : -1.0) : 0.0);
      }
      if (kind == 2 && v[0] > 3.0) return 1.0 / 0.0;
    }
    return fx;
  };
  int ret = lbfgs_optimize(xx, *f, fn, *param);
  std::memcpy(x, xx.data(), n * sizeof(double));
  if (evals) *evals = ev;
  return ret;
}

// ---- pieces exposed only so that tests/test_ref_pins.py can compare them with the reference-compiled fragments ----
// BandedSystem (minco.hpp:43-198) on a dense row-major N x N matrix, rhs b [N][2]; mode 0 solve, 1 solveAdj.
int orc_banded(int N, int p, int q, const double* A, double* b, int mode, double* band_out) {
  BandedSystem bs;
  bs.create(N, p, q);
  for (int i = 0; i < N; i++)
    for (int j = std::max(0, i - p); j <= std::min(N - 1, i + q); j++) bs(i, j) = A[(size_t)i * N + j];
  bs.factorizeLU();
  if (mode == 0) bs.solve(b);
  else bs.solveAdj(b);
  if (band_out) std::memcpy(band_out, bs.ptrData.data(), bs.ptrData.size() * sizeof(double));
  return 0;
}
// which: 0 RealT2VirtualT, 1 VirtualT2RealT, 2 backwardGradT (optimizer.cpp:573-591, 1088-1106)
void orc_tmaps(int n, const double* in, int which, const double* gradT, double* out) {
  Vec a(in, in + n), o;
  if (which == 0) MSPlanner::RealT2VirtualT(a, out);
  else if (which == 1) { MSPlanner::VirtualT2RealT(in, n, o); std::memcpy(out, o.data(), n * sizeof(double)); }
  else { Vec g(gradT, gradT + n); MSPlanner::backwardGradT(in, g, out); }
}
void orc_smoothed_l1(double eps, double x, double* f, double* df) {   // optimizer.cpp:1069-1086
  MSPlanner pl;
  pl.p.smoothEps = eps;
  pl.positiveSmoothedL1(x, *f, *df);
}

// The reference's yaml defaults (back_end/config/global_planning3ms.yaml, plan_tester/config/car3ms.yaml,
// plan_tester/launch/planner_sim.launch:41-46; lbfgs.hpp:76-128), stated here a second time so that the CPU arm of
// bench.py needs nothing of the product library; tests/test_capi_cpu.py asserts the two agree byte for byte.
static void orc_lbfgs_defaults(alore_lbfgs_params_t* l) {
  l->mem_size = 8; l->past = 3; l->max_iterations = 0; l->max_linesearch = 64;
  l->g_epsilon = 1.0e-5; l->delta = 1.0e-6; l->min_step = 1.0e-20; l->max_step = 1.0e+20;
  l->f_dec_coeff = 1.0e-4; l->s_curv_coeff = 0.9; l->cautious_factor = 1.0e-6; l->machine_prec = 1.0e-16;
}
void orc_params_default(alore_params_t* p) {
  std::memset(p, 0, sizeof(*p));
  p->max_vel = 3.0; p->min_vel = -3.0; p->max_acc = 2.0; p->max_omega = 3.0; p->max_domega = 4.0;
  p->max_centripetal_acc = 50.0; p->if_directly_constrain_v_omega = 0; p->if_standard_diff = 1;
  p->ICR[0] = 0.3; p->ICR[1] = -0.3; p->ICR[2] = 0.2;
  p->mean_time_lowBound = 0.5; p->mean_time_uppBound = 2.0;
  p->smoothEps = 0.01; p->safeDis = 0.6; p->finalMinSafeDis = 0.10; p->finalSafeDisCheckNum = 16; p->safeReplanMaxTime = 3;
  p->pw_time = 50; p->pw_acc = 300; p->pw_domega = 300; p->pw_collision = 500000; p->pw_moment = 300; p->pw_mean_time = 300; p->pw_cen_acc = 300;
  p->ppw_time = 20; p->ppw_bigpath_sdf = 200000; p->ppw_mean_time = 100; p->ppw_moment = 1000; p->ppw_acc = 100; p->ppw_domega = 100;
  p->energyWeights[0] = 0.33; p->energyWeights[1] = 1.0;
  for (int i = 0; i < 2; i++) {
    p->EqualLambda[i] = 0; p->EqualRho[i] = 10000.0; p->EqualRhoMax[i] = 1.0e10; p->EqualGamma[i] = 9.0;
    p->CutEqualLambda[i] = 0; p->CutEqualRho[i] = 1000.0; p->CutEqualRhoMax[i] = 1.0e10; p->CutEqualGamma[i] = 5.0;
  }
  p->EqualTolerance[0] = 0.01; p->EqualTolerance[1] = 0.0; p->CutEqualTolerance[0] = 0.5; p->CutEqualTolerance[1] = 0.0;
  orc_lbfgs_defaults(&p->path_lbfgs);
  p->path_lbfgs.mem_size = 256; p->path_lbfgs.past = 2; p->path_lbfgs.g_epsilon = 0.0; p->path_lbfgs.min_step = 0.0;
  p->path_lbfgs.delta = 5.0e-2; p->path_lbfgs.max_iterations = 8000;
  p->normal_past = 2; p->shot_path_past = 8; p->shot_path_horizon = 0.5;
  orc_lbfgs_defaults(&p->lbfgs);
  p->lbfgs.mem_size = 256; p->lbfgs.past = 3; p->lbfgs.g_epsilon = 0.0; p->lbfgs.min_step = 1.0e-32; p->lbfgs.delta = 5.0e-4;
  p->lbfgs.max_iterations = 8000;
  p->sparseResolution = 8; p->n_checkpoints = 1; p->check_point[0][0] = 0.0; p->check_point[0][1] = 0.0;
  p->alm_max_outer = 0;
}

void orc_set_exact_chain_weights(int on) { orc::g_exact_chain_weights = on != 0; }

void orc_set_trig_portable(int on) { orc::g_trig_portable = on != 0; }
void orc_sincos(double x, int portable, double* s, double* c) {
  if (portable) orc::ptrig::sincos(x, *s, *c);
  else { *s = std::sin(x); *c = std::cos(x); }
}

int orc_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
