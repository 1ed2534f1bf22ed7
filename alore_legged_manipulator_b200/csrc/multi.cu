// multi.cu — one library object driving 1..8 GPUs of one box (SURVEY.md section 8b "Ownership/Threading", 8e).
//
// The reference's caller is C++ (PlanManager owns ONE shared_ptr<SDFmap> and ONE shared_ptr<MSPlanner>,
// planning_ddr_opt/plan_manager/include/plan_manager/plan_manager.hpp:120-123, and calls minco_plan on the ROS thread,
// :662-670); it cannot start torchrun.  alore_multi gives it the multi-GPU path behind the same two calls:
//   alore_multi_esdf_update   the occupancy grid goes to every device, every device rebuilds the ESDF locally
//                             (1 B/cell up, cheaper than broadcasting 8 B/cell); the host mirror is written once
//   alore_multi_opt_batch     the candidate array is cut into contiguous, cost-balanced blocks (sum of pieces), one host
//                             thread per GPU runs alore_opt_batch on its block, every GPU reduces its block to
//                             (best cost, global index) and the pairs are ALL-GATHERED OVER NCCL (the only exchange of
//                             the path); the host reads the gathered pairs and returns the winner
// NCCL is loaded with dlopen (no link-time dependency, no clash with another NCCL copy in the process).
#include <dlfcn.h>

#include <thread>

#include "common.cuh"

namespace {

// the handful of NCCL entry points used, resolved at run time
typedef struct ncclComm* ncclComm_t;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclChar = 0 } ncclDataType_t;
struct Nccl {
  void* h = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load(std::string& err) {
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (h) break;
    }
    if (!h) { err = std::string("cannot load NCCL: ") + dlerror(); return false; }
#define SYM(f, n) f = reinterpret_cast<decltype(f)>(dlsym(h, n)); if (!f) { err = std::string("NCCL lacks ") + n; return false; }
    SYM(CommInitAll, "ncclCommInitAll") SYM(CommDestroy, "ncclCommDestroy") SYM(AllGather, "ncclAllGather")
    SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    return true;
  }
};

struct BestPair { double cost; long long idx; };   // 16 bytes per rank on the wire

}  // namespace

struct alore_multi {
  int n = 0;
  std::vector<int> devices;
  std::vector<alore_ctx*> ctx;
  std::vector<ncclComm_t> comm;
  std::vector<BestPair*> d_mine;     // per device: this rank's pair
  std::vector<BestPair*> d_all;      // per device: n gathered pairs
  Nccl nccl;
  bool have_nccl = false;
  std::string err;
};

static int multi_fail(alore_multi* m, int code, const std::string& msg) {
  if (m) m->err = msg;
  return code;
}

extern "C" {

int alore_create_multi(const int* devices, int n, alore_multi** out) {
  if (!out || !devices || n < 1 || n > 64) return ALORE_EINVAL;
  *out = nullptr;
  alore_multi* m = new alore_multi();
  m->n = n;
  m->devices.assign(devices, devices + n);
  m->ctx.assign(n, nullptr);
  m->d_mine.assign(n, nullptr);
  m->d_all.assign(n, nullptr);
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < i; j++)
      if (devices[j] == devices[i]) { alore_destroy_multi(m); return ALORE_EINVAL; }
    const int rc = alore_create(devices[i], &m->ctx[i]);
    if (rc) { alore_destroy_multi(m); return rc; }
    cudaSetDevice(devices[i]);
    if (cudaMalloc(&m->d_mine[i], sizeof(BestPair)) != cudaSuccess || cudaMalloc(&m->d_all[i], n * sizeof(BestPair)) != cudaSuccess) {
      alore_destroy_multi(m);
      return ALORE_ENOMEM;
    }
  }
  if (n > 1) {                       // one communicator per device, one process (ncclCommInitAll)
    std::string e;
    if (!m->nccl.load(e)) { alore_destroy_multi(m); return ALORE_ECUDA; }
    m->comm.assign(n, nullptr);
    const ncclResult_t r = m->nccl.CommInitAll(m->comm.data(), n, m->devices.data());
    if (r != ncclSuccess) { m->comm.clear(); alore_destroy_multi(m); return ALORE_ECUDA; }
    m->have_nccl = true;
  }
  *out = m;
  return ALORE_OK;
}

void alore_destroy_multi(alore_multi* m) {
  if (!m) return;
  for (int i = 0; i < m->n; i++) {
    if ((int)m->comm.size() > i && m->comm[i]) m->nccl.CommDestroy(m->comm[i]);
    if (m->ctx[i]) {
      cudaSetDevice(m->devices[i]);
      if (m->d_mine[i]) cudaFree(m->d_mine[i]);
      if (m->d_all[i]) cudaFree(m->d_all[i]);
      alore_destroy(m->ctx[i]);
    }
  }
  if (m->nccl.h) dlclose(m->nccl.h);
  delete m;
}

int alore_multi_size(const alore_multi* m) { return m ? m->n : 0; }
alore_ctx* alore_multi_ctx(alore_multi* m, int i) { return (m && i >= 0 && i < m->n) ? m->ctx[i] : nullptr; }
const char* alore_multi_last_error(const alore_multi* m) { return m ? m->err.c_str() : ""; }

int alore_multi_esdf_update(alore_multi* m, const alore_map_geom_t* geom, const uint8_t* occ, int min_x, int min_y, int max_x, int max_y,
                            double* dist_inout, int ref_compat) {
  if (!m || !geom || !occ || !dist_inout) return ALORE_EINVAL;
  std::vector<int> rc(m->n, ALORE_OK);
  std::vector<std::thread> th;
  for (int i = 0; i < m->n; i++)
    th.emplace_back([&, i]() {
      // device 0 also writes the host mirror (SDFmap::distance_buffer_all_); the replicas keep their result in HBM only
      rc[i] = alore_esdf_update_impl(m->ctx[i], geom, occ, min_x, min_y, max_x, max_y, i == 0 ? dist_inout : nullptr, ref_compat);
    });
  for (auto& t : th) t.join();
  for (int i = 0; i < m->n; i++)
    if (rc[i]) return multi_fail(m, rc[i], std::string("device ") + std::to_string(m->devices[i]) + ": " + alore_last_error(m->ctx[i]));
  return ALORE_OK;
}

// Contiguous blocks with (nearly) equal sums of pieces: offs[0..n].
static void balanced_blocks(const int32_t* piece_off, int B, int n, std::vector<int>& offs) {
  offs.assign(n + 1, 0);
  const long long tot = piece_off[B];
  int b = 0;
  for (int r = 1; r < n; r++) {
    const long long target = tot * r / n;
    while (b < B && piece_off[b] < target) b++;
    offs[r] = std::max(b, offs[r - 1]);
  }
  offs[n] = B;
}

int alore_multi_opt_batch(alore_multi* m, const alore_params_t* prm, const alore_candidates_t* c, alore_results_t* out, double* best_cost,
                          int32_t* best_idx, int32_t* block_offsets /* [n+1], may be NULL */) {
  if (!m || !prm || !c || !out || c->B <= 0 || !c->piece_off) return ALORE_EINVAL;
  const int n = m->n, B = c->B;
  std::vector<int> offs;
  balanced_blocks(c->piece_off, B, n, offs);
  if (block_offsets) for (int i = 0; i <= n; i++) block_offsets[i] = offs[i];
  std::vector<int> rc(n, ALORE_OK);
  std::vector<BestPair> mine(n, BestPair{1.7976931348623157e308, -1});
  std::vector<std::thread> th;
  for (int i = 0; i < n; i++)
    th.emplace_back([&, i]() {
      const int b0 = offs[i], b1 = offs[i + 1], nb = b1 - b0;
      if (nb <= 0) return;
      const int p0 = c->piece_off[b0];
      std::vector<int32_t> po(nb + 1);
      for (int k = 0; k <= nb; k++) po[k] = c->piece_off[b0 + k] - p0;
      alore_candidates_t s = *c;
      s.B = nb;
      s.piece_off = po.data();
      s.inner_pts = c->inner_pts + 2 * (size_t)(p0 - b0);
      s.init_T = c->init_T + b0;
      s.inner_init_pos = c->inner_init_pos + 3 * (size_t)p0;
      s.start_state = c->start_state + 6 * (size_t)b0;
      s.final_state = c->final_state + 6 * (size_t)b0;
      s.start_xytheta = c->start_xytheta + 3 * (size_t)b0;
      s.final_xytheta = c->final_xytheta + 3 * (size_t)b0;
      s.if_cut = c->if_cut + b0;
      alore_results_t r = *out;
      r.ok = out->ok + b0; r.status = out->status + b0; r.replans = out->replans + b0; r.alm_iters = out->alm_iters + b0;
      r.evals = out->evals + b0; r.cost = out->cost + b0; r.tail_s = out->tail_s + b0;
      r.inner_pts = out->inner_pts + 2 * (size_t)(p0 - b0);
      r.piece_T = out->piece_T + p0;
      r.coeffs = out->coeffs + 12 * (size_t)p0;
      rc[i] = alore_opt_batch(m->ctx[i], prm, &s, &r);
      if (rc[i]) return;
      for (int k = 0; k < nb; k++)      // lowest final cost among the successful candidates of the block; ties -> lowest index
        if (r.ok[k] == 1 && r.cost[k] < mine[i].cost) { mine[i].cost = r.cost[k]; mine[i].idx = b0 + k; }
    });
  for (auto& t : th) t.join();
  for (int i = 0; i < n; i++)
    if (rc[i]) return multi_fail(m, rc[i], std::string("device ") + std::to_string(m->devices[i]) + ": " + alore_last_error(m->ctx[i]));
  // the one exchange of the path: all-gather of (best cost, global index), 16 bytes per rank, over NCCL
  std::vector<BestPair> all(n);
  if (n == 1) {
    all[0] = mine[0];
  } else {
    for (int i = 0; i < n; i++) {
      cudaSetDevice(m->devices[i]);
      cudaMemcpyAsync(m->d_mine[i], &mine[i], sizeof(BestPair), cudaMemcpyHostToDevice, m->ctx[i]->stream);
    }
    if (m->nccl.GroupStart() != ncclSuccess) return multi_fail(m, ALORE_ECUDA, "ncclGroupStart failed");
    for (int i = 0; i < n; i++) {
      const ncclResult_t r = m->nccl.AllGather(m->d_mine[i], m->d_all[i], sizeof(BestPair), ncclChar, m->comm[i], m->ctx[i]->stream);
      if (r != ncclSuccess) { m->nccl.GroupEnd(); return multi_fail(m, ALORE_ECUDA, std::string("ncclAllGather: ") + m->nccl.GetErrorString(r)); }
    }
    if (m->nccl.GroupEnd() != ncclSuccess) return multi_fail(m, ALORE_ECUDA, "ncclGroupEnd failed");
    cudaSetDevice(m->devices[0]);
    cudaMemcpyAsync(all.data(), m->d_all[0], n * sizeof(BestPair), cudaMemcpyDeviceToHost, m->ctx[0]->stream);
    for (int i = 0; i < n; i++) {
      cudaSetDevice(m->devices[i]);
      const cudaError_t e = cudaStreamSynchronize(m->ctx[i]->stream);
      if (e != cudaSuccess) return multi_fail(m, ALORE_ECUDA, std::string("gather: ") + cudaGetErrorString(e));
    }
  }
  double bc = 1.7976931348623157e308;
  long long bi = -1;
  for (int i = 0; i < n; i++)
    if (all[i].idx >= 0 && (all[i].cost < bc || (all[i].cost == bc && all[i].idx < bi))) { bc = all[i].cost; bi = all[i].idx; }
  if (best_cost) *best_cost = bc;
  if (best_idx) *best_idx = (int32_t)bi;
  return ALORE_OK;
}

}  // extern "C"
