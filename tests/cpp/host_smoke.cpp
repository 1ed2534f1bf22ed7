// C++ host-side smoke: the drop-in classes of alore_host.hpp against libalore_b200.so.
// Prints one line per check; exit code 0 on success.  Run on the GPU box by tests/test_cpp_host.py.
#include <cstdio>
#include <cstdlib>

#include "../../alore_legged_manipulator_b200/csrc/host/alore_host.hpp"

int main() {
  alore::Context ctx(0);
  alore::SDFmap map(ctx, 0.05, 100.0, -5.0, 5.0 - 0.025, -5.0, 5.0 - 0.025);
  if (map.GLX_SIZE_ != 200 || map.GLY_SIZE_ != 200) { std::printf("bad geometry %d %d\n", map.GLX_SIZE_, map.GLY_SIZE_); return 1; }
  for (int x = 0; x < 200; x++)
    for (int y = 0; y < 200; y++) map.gridmap()[x * 200 + y] = (x == 0 || y == 0 || x == 199 || y == 199) ? alore::SDFmap::Occupied : alore::SDFmap::Unoccupied;
  for (double t = -1.0; t <= 1.0; t += 0.02) map.setObs({t, 1.5});     // a wall segment painted through the reference's setter
  map.forceUpdateESDF();
  const double d0 = map.getDistanceReal({0.0, 0.0});
  std::printf("dist(0,0)=%.4f\n", d0);
  if (!(d0 > 1.4 && d0 < 1.6)) return 2;
  alore::Vec2d g;
  const double d1 = map.getDistWithGradBilinear({0.0, 1.0}, g);
  std::printf("bilinear(0,1)=%.4f grad=(%.3f,%.3f)\n", d1, g[0], g[1]);
  if (!(g[1] < -0.9)) return 3;

  alore::MSPlanner planner(ctx, map);
  planner.params.alm_max_outer = 20;
  alore::FlatTrajData ft;                       // a short straight leg, 4 pieces
  const int N = 4;
  const double L = 2.4, T0 = 0.5;
  for (int i = 1; i < N; i++) {
    ft.UnOccupied_traj_pts.push_back({0.0, L * i / N, T0 * i});
    ft.UnOccupied_positions.push_back({-3.0 + L * i / N, -3.0, 0.0});
  }
  ft.UnOccupied_initT = T0;
  ft.final_state[1][0] = L;
  ft.start_state_XYTheta = {-3.0, -3.0, 0.0};
  ft.final_state_XYTheta = {-3.0 + L, -3.0, 0.0};
  const bool ok = planner.minco_plan(ft);
  const alore::PlanResult& r = planner.final_result();
  std::printf("minco_plan ok=%d status=%d evals=%d cost=%.6f T0=%.4f\n", (int)ok, r.status, r.evals, r.cost, r.piece_T[0]);
  if (!ok || r.piece_T.size() != (size_t)N || r.coeffs.size() != (size_t)12 * N) return 4;
  std::printf("HOST_SMOKE_OK\n");
  return 0;
}
