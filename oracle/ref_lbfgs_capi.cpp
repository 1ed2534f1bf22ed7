// ref_lbfgs_capi.cpp — builds oracle/_ref/libref_lbfgs.so from the REFERENCE'S OWN lbfgs.hpp, included from
// where it lies (/root/reference/planning_ddr_opt/back_end/include/gcopter/lbfgs.hpp; never copied), through
// the Eigen stand-in in oracle/eigen_shim.  Test infrastructure only: pins the oracle's lbfgs restatement.
#include <cstring>

#include "gcopter/lbfgs.hpp"

#include "../include/alore_b200.h"

namespace {
struct Problem {
  int kind;   // 0: extended Rosenbrock, 1: ill-conditioned quadratic with a kink (nonsmooth), 2: returns inf when far
  int evals;
};
double eval_cb(void* inst, const Eigen::VectorXd& x, Eigen::VectorXd& g) {
  Problem* p = static_cast<Problem*>(inst);
  p->evals++;
  const int n = (int)x.size();
  double fx = 0.0;
  if (p->kind == 0) {
    for (int i = 0; i < n; i += 2) {
      double t1 = 1.0 - x(i);
      double t2 = 10.0 * (x(i + 1) - x(i) * x(i));
      g(i + 1) = 20.0 * t2;
      g(i) = -2.0 * (x(i) * g(i + 1) + t1);
      fx += t1 * t1 + t2 * t2;
    }
  } else {
    for (int i = 0; i < n; i++) {
      const double w = 1.0 + 50.0 * i;
      const double a = x(i) - 0.1 * i;
      fx += 0.5 * w * a * a + (p->kind == 1 ? std::fabs(x(i)) : 0.0);
      g(i) = w * a + (p->kind == 1 ? (x(i) > 0 ? 1.0 : -1.0) : 0.0);
    }
    if (p->kind == 2 && x(0) > 3.0) return 1.0 / 0.0;
  }
  return fx;
}
}  // namespace

extern "C" int ref_lbfgs_run(int kind, int n, double* x, const alore_lbfgs_params_t* prm, double* f, int* evals) {
  lbfgs::lbfgs_parameter_t p;
  p.mem_size = prm->mem_size; p.g_epsilon = prm->g_epsilon; p.past = prm->past; p.delta = prm->delta;
  p.max_iterations = prm->max_iterations; p.max_linesearch = prm->max_linesearch; p.min_step = prm->min_step;
  p.max_step = prm->max_step; p.f_dec_coeff = prm->f_dec_coeff; p.s_curv_coeff = prm->s_curv_coeff;
  p.cautious_factor = prm->cautious_factor; p.machine_prec = prm->machine_prec;
  Eigen::VectorXd xv(n);
  for (int i = 0; i < n; i++) xv(i) = x[i];
  Problem pb{kind, 0};
  double fx = 0.0;
  int ret = lbfgs::lbfgs_optimize(xv, fx, eval_cb, nullptr, nullptr, &pb, p);
  for (int i = 0; i < n; i++) x[i] = xv(i);
  *f = fx;
  if (evals) *evals = pb.evals;
  return ret;
}
