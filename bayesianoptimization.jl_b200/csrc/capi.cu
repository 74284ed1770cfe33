// capi.cu -- the C ABI of libb200bo.so (include/b200bo.h).  Host orchestration only; all arithmetic is in the
// CUDA kernels (kmat.cu, chol.cu, solve.cu, acq.cu, mll.cu).  No CPU fallback: without a device every compute
// entry fails with B200BO_ERR_CUDA.
#include <cstdlib>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <limits>
#include <chrono>
#include <new>
#include <thread>
#include "common.cuh"
#include "lbfgs.cuh"
#include "direct.h"
#include "handle.h"

using namespace b200bo;

namespace {

thread_local std::string g_err;   // for errors raised without a handle

int32_t fail(b200bo_handle_t h, int32_t code, const std::string& msg) {
  if (h) h->err = msg; else g_err = msg;
  return code;
}
#define CU(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e__ = (call);                                                                            \
    if (e__ != cudaSuccess)                                                                              \
      return fail(h, B200BO_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));              \
  } while (0)

int fam_of(int kind) { return kind >> 1; }          // SE, Mat12, Mat32, Mat52
bool iso_of(int kind) { return (kind & 1) == 0; }

int num_params(const b200bo_handle_s* h) {
  return 1 + (h->mean_kind == B200BO_MEAN_CONST ? 1 : 0) + (int)h->hp.ll.size() + 1;
}

void free_device(b200bo_handle_s* h) {
  release_cholesky_graph(h);
  cudaFree(h->dX); cudaFree(h->dZ); cudaFree(h->dZk); cudaFree(h->dy); cudaFree(h->dw); cudaFree(h->dalpha); cudaFree(h->dz); cudaFree(h->dflags); cudaFree(h->dinv_ell);
  cudaFree(h->dL); cudaFree(h->dLinv); cudaFree(h->dLinvT); cudaFree(h->dV); cudaFree(h->dKi); cudaFree(h->dWT); cudaFree(h->dTT); cudaFree(h->dSl); cudaFree(h->dSe); cudaFree(h->dscal); cudaFree(h->dinfo);
  cudaFree(h->dcta_best); cudaFree(h->dbest); cudaFree(h->dpart); cudaFree(h->dD); h->dD = nullptr;
  cudaFree(h->dWs); cudaFree(h->dKs); cudaFree(h->dBs); cudaFree(h->dWe); cudaFree(h->dKe); cudaFree(h->dMuP); cudaFree(h->dWg); cudaFree(h->dcta_best2);
  h->dWs = h->dKs = h->dBs = nullptr; h->dWe = h->dKe = h->dMuP = h->dWg = nullptr; h->dcta_best2 = nullptr;
  h->bs_bytes = h->part_bytes = h->wg_bytes = 0; h->bs_np = h->bs_ch = h->nbest2 = 0; h->acq_ready = 0; h->bs_grad = false; h->wt_valid = false;
  h->dz = nullptr; h->dflags = nullptr; h->dKi = h->dWT = h->dTT = nullptr; h->dSl = nullptr; h->dSe = nullptr;
  h->dX = h->dZ = h->dZk = h->dy = h->dw = h->dalpha = h->dinv_ell = h->dL = h->dLinv = h->dLinvT = h->dV = h->dscal = h->dpart = nullptr;
  h->dinfo = nullptr; h->dcta_best = nullptr; h->dbest = nullptr;
}

int32_t sync_inv_ell(b200bo_handle_s* h);

int32_t alloc_device(b200bo_handle_s* h, int64_t cap) {
  cap = std::max<int64_t>(NB, (cap + NB - 1) / NB * NB);
  free_device(h);
  h->cap = cap; h->ld = cap;
  h->nslots = h->lite ? 1 : std::max<int64_t>(h->num_sms, cap / TILE_N);
  const int64_t nb = cap / NB;
  const int64_t T = cap / 64;
  CU(cudaMalloc(&h->dX, sizeof(double) * cap * h->D));
  CU(cudaMalloc(&h->dZ, sizeof(double) * cap * h->D));
  CU(cudaMalloc(&h->dZk, sizeof(double) * (cap / 64) * (8 * ((h->D + 3) / 4) + 2) * 64));
  CU(cudaMalloc(&h->dy, sizeof(double) * cap));
  CU(cudaMalloc(&h->dw, sizeof(double) * cap));
  CU(cudaMalloc(&h->dalpha, sizeof(double) * cap));
  CU(cudaMalloc(&h->dz, sizeof(double) * cap));
  CU(cudaMalloc(&h->dflags, sizeof(int) * (cap / NB)));
  CU(cudaMemsetAsync(h->dflags, 0, sizeof(int) * (cap / NB), h->stream));
  h->solve_epoch = 0;
  CU(cudaMalloc(&h->dinv_ell, sizeof(double) * 2 * h->D));
  CU(cudaMalloc(&h->dL, sizeof(double) * cap * cap));
  CU(cudaMalloc(&h->dLinv, sizeof(double) * nb * NB * NB));
  CU(cudaMalloc(&h->dLinvT, sizeof(double) * nb * NB * NB));
  CU(cudaMalloc(&h->dV, sizeof(double) * h->nslots * TILE_N * cap));
  CU(cudaMalloc(&h->dscal, sizeof(double) * 64));
  CU(cudaMalloc(&h->dinfo, 4 * sizeof(int)));
  CU(cudaMemsetAsync(h->dinfo, 0, 4 * sizeof(int), h->stream));
  CU(cudaMalloc(&h->dcta_best, sizeof(b200bo_best_t) * h->nslots));
  CU(cudaMalloc(&h->dbest, sizeof(b200bo_best_t)));
  CU(cudaMalloc(&h->dpart, sizeof(double) * T * T * 35));
  CU(cudaMemsetAsync(h->dX, 0, sizeof(double) * cap * h->D, h->stream));
  CU(cudaMemsetAsync(h->dZ, 0, sizeof(double) * cap * h->D, h->stream));
  CU(cudaMemsetAsync(h->dZk, 0, sizeof(double) * (cap / 64) * (8 * ((h->D + 3) / 4) + 2) * 64, h->stream));
  CU(cudaMemsetAsync(h->dLinv, 0, sizeof(double) * nb * NB * NB, h->stream));     // K2 writes one triangle of each block only
  CU(cudaMemsetAsync(h->dLinvT, 0, sizeof(double) * nb * NB * NB, h->stream));
  CU(cudaMemsetAsync(h->dalpha, 0, sizeof(double) * cap, h->stream));
  CU(cudaMemsetAsync(h->dy, 0, sizeof(double) * cap, h->stream));
  CU(make_tensor_maps(h));
  return sync_inv_ell(h);
}

int32_t ensure_io(b200bo_handle_s* h, int64_t bytes) {
  if (bytes <= h->dio_bytes) return B200BO_OK;
  if (h->dio) cudaFree(h->dio);
  h->dio = nullptr; h->dio_bytes = 0;
  CU(cudaMalloc(&h->dio, (size_t)bytes));
  h->dio_bytes = bytes;
  return B200BO_OK;
}

// [inv_ell (D) | centre (D)]: the centre (mid-range of the observations, in x units) only shifts the inputs of K1's Gram-form
// distances (kmat.cu), whose rounding error grows with |z|^2; every other kernel works with coordinate differences
void upload_inv_ell(b200bo_handle_s* h, std::vector<double>& ie) {
  const int D = h->D;
  ie.assign(2 * D, 0.0);
  for (int d = 0; d < D; ++d) ie[d] = exp(-(h->iso ? h->hp.ll[0] : h->hp.ll[d]));
  const int64_t N = std::min<int64_t>(h->N, (int64_t)(h->hX.size() / D));
  for (int d = 0; d < D && N > 0; ++d) {
    double lo = h->hX[d], hi = lo;
    for (int64_t i = 1; i < N; ++i) { const double x = h->hX[i * D + d]; lo = x < lo ? x : lo; hi = x > hi ? x : hi; }
    const double c = 0.5 * (lo + hi);
    ie[D + d] = std::isfinite(c) ? c : 0.0;
  }
}

// dinv_ell always holds the inverse length-scales of the current parameters (an empty model evaluates the prior through K6, which
// multiplies its zero gradient accumulators by them: they must be finite from the first call on)
int32_t sync_inv_ell(b200bo_handle_s* h) {
  if (!h->dinv_ell) return B200BO_OK;
  std::vector<double> ie;
  upload_inv_ell(h, ie);
  CU(cudaMemcpyAsync(h->dinv_ell, ie.data(), sizeof(double) * 2 * h->D, cudaMemcpyHostToDevice, h->stream));   // pageable source: staged before return
  return B200BO_OK;
}

// Sigma assembly + Cholesky (with the make_posdef! jitter rule) + alpha + mll on the current data and params.
int32_t upload_data(b200bo_handle_s* h);

int32_t refit(b200bo_handle_s* h) {
  h->jitter = 0;
  h->acq_ready = 0;                     // W = L^-1 and its int8 slices (acq_i8.cu) belong to the previous factor
  h->wt_valid = false;
  if (h->need_upload) { const int32_t rc = upload_data(h); if (rc) return rc; }
  if (h->N == 0) { h->fitted = true; h->mll = 0.0; h->Np = 0; return B200BO_OK; }
  std::vector<double> ie;
  upload_inv_ell(h, ie);
  CU(cudaMemcpyAsync(h->dinv_ell, ie.data(), sizeof(double) * 2 * h->D, cudaMemcpyHostToDevice, h->stream));
  CU(launch_scale_inputs(h, 0, h->Np));
  const double sf2 = exp(2.0 * h->hp.lsigma);
  double noise = exp(2.0 * h->hp.lognoise) + std::numeric_limits<double>::epsilon();   // quirk 10
  for (;;) {
    CU(cudaEventRecord(h->ev[0], h->stream));
    CU(launch_kmat(h, h->dL, h->ld, h->N, h->Np, noise, true));
    CU(cudaEventRecord(h->ev[1], h->stream));
    CU(launch_cholesky(h));
    CU(cudaEventRecord(h->ev[2], h->stream));
    int info = 0;
    CU(cudaMemcpyAsync(&info, h->dinfo, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (info == 0) break;
    if (h->jitter >= 10) return fail(h, B200BO_ERR_NOTPD, "covariance not positive definite after 10 jitter retries");
    noise += 1e-6 * (sf2 + noise);      // 1e-6 * tr(Sigma)/n with diag(Sigma) = sf2 + noise (EXT make_posdef!)
    h->jitter++;
  }
  h->noise_total = noise;
  CU(launch_alpha_mll(h, true));      // z = L^-1 (y - m) came out of the factorisation
  CU(cudaEventRecord(h->ev[3], h->stream));
  double sc[2];
  CU(cudaMemcpyAsync(sc, h->dscal, sizeof(sc), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  h->mll = -0.5 * (sc[1] + sc[0] + (double)h->N * 1.8378770664093453);
  cudaEventElapsedTime(&h->timing[B200BO_T_KMAT], h->ev[0], h->ev[1]);
  cudaEventElapsedTime(&h->timing[B200BO_T_CHOL], h->ev[1], h->ev[2]);
  cudaEventElapsedTime(&h->timing[B200BO_T_ALPHA], h->ev[2], h->ev[3]);
  float syrk = 0.f;
  for (int i = 0; i + 1 < h->syrk_ev_used; i += 2) { float t = 0.f; cudaEventElapsedTime(&t, h->syrk_ev[i], h->syrk_ev[i + 1]); syrk += t; }
  h->timing[B200BO_T_SYRK] = syrk;
  h->fitted = true;
  return B200BO_OK;
}

int32_t ensure_fitted(b200bo_handle_s* h) { return h->fitted ? B200BO_OK : refit(h); }

int32_t upload_data(b200bo_handle_s* h) {
  const int64_t N = h->N;
  if (N > h->cap) {
    int32_t rc = alloc_device(h, std::max<int64_t>(N, 2 * h->cap));
    if (rc) return rc;
  }
  h->Np = (N + NB - 1) / NB * NB;
  if (N > 0) {
    CU(cudaMemcpyAsync(h->dX, h->hX.data(), sizeof(double) * N * h->D, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->dy, h->hy.data(), sizeof(double) * N, cudaMemcpyHostToDevice, h->stream));
    if (h->Np > N) {
      CU(cudaMemsetAsync(h->dX + N * h->D, 0, sizeof(double) * (h->Np - N) * h->D, h->stream));
      CU(cudaMemsetAsync(h->dy + N, 0, sizeof(double) * (h->Np - N), h->stream));
    }
  }
  h->fitted = false;
  h->need_upload = false;
  return B200BO_OK;
}

int32_t check_acq(b200bo_handle_t h, int32_t kind, int32_t n_params, bool grad) {
  static const int need[6] = {1, 1, 1, 0, 2, 0};
  if (kind < 0 || kind > B200BO_ACQ_MAXMEAN) return fail(h, B200BO_ERR_ARG, "unknown acquisition kind");
  if (n_params < need[kind]) return fail(h, B200BO_ERR_ARG, "too few acquisition parameters");
  if (grad && kind == B200BO_ACQ_TS)
    return fail(h, B200BO_ERR_ARG, "ThompsonSamplingSimple is derivative-free (reference src/acquisition.jl:7-9)");
  return B200BO_OK;
}

// While a multi handle fans a call out to its replicas, each replica -- the parent included -- must behave as a plain single-GPU
// handle: no recursion into the fan-out and NO communicator exchange of its own (a lone rank entering a collective never returns).
struct SoloScope {
  b200bo_handle_s* h;
  std::vector<b200bo_handle_s*> kids;
  bool was;
  explicit SoloScope(b200bo_handle_s* r) : h(r), was(r->in_multi) { kids.swap(h->replicas); h->in_multi = true; }
  ~SoloScope() { kids.swap(h->replicas); h->in_multi = was; }
};

// every replica of a multi handle: the parent first
std::vector<b200bo_handle_s*> all_replicas(b200bo_handle_s* h) {
  std::vector<b200bo_handle_s*> v{h};
  v.insert(v.end(), h->replicas.begin(), h->replicas.end());
  return v;
}

// run f(replica, rank) on every replica, the children on their own host threads (the entries block on their streams)
template <class F>
int32_t on_replicas(b200bo_handle_s* h, F f) {
  const std::vector<b200bo_handle_s*> reps = all_replicas(h);
  std::vector<int32_t> rc(reps.size(), B200BO_OK);
  std::vector<std::thread> th;
  for (size_t r = 1; r < reps.size(); ++r) th.emplace_back([&, r] { rc[r] = f(reps[r], (int)r); });
  rc[0] = f(reps[0], 0);
  for (auto& t : th) t.join();
  for (size_t r = 0; r < reps.size(); ++r)
    if (rc[r] != B200BO_OK) { if (r > 0) h->err = "replica " + std::to_string(r) + ": " + reps[r]->err; return rc[r]; }
  return B200BO_OK;
}

// a "lite" model on the parent's device (no acquisition solve panels, own streams and buffers): MAP sweeps and the joint posterior sample
int32_t new_worker(b200bo_handle_s* h, int64_t cap, b200bo_handle_s** out) {
  b200bo_handle_s* wk = new (std::nothrow) b200bo_handle_s();
  if (!wk) return fail(h, B200BO_ERR_ALLOC, "out of host memory");
  wk->device = h->device; wk->D = h->D; wk->kernel_kind = h->kernel_kind; wk->mean_kind = h->mean_kind; wk->fam = h->fam; wk->iso = h->iso;
  wk->num_sms = h->num_sms; wk->lite = true; wk->sweep_workers = 0; wk->hp = h->hp;
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  if (cudaStreamCreateWithPriority(&wk->stream, cudaStreamNonBlocking, hi) != cudaSuccess ||
      cudaStreamCreateWithPriority(&wk->stream2, cudaStreamNonBlocking, lo) != cudaSuccess) { delete wk; return fail(h, B200BO_ERR_CUDA, "worker streams"); }
  for (auto& e : wk->ev) cudaEventCreate(&e);
  const int32_t rcw = alloc_device(wk, cap);
  if (rcw != B200BO_OK) { h->err = wk->err; b200bo_destroy(wk); return rcw; }
  *out = wk;
  return B200BO_OK;
}

}  // namespace

extern "C" {

B200BO_API int32_t b200bo_version(void) { return 100; }

B200BO_API const char* b200bo_last_error(b200bo_handle_t h) { return h ? h->err.c_str() : g_err.c_str(); }

B200BO_API int32_t b200bo_create(b200bo_handle_t* out, int32_t device, int32_t D, int64_t capacity, int32_t kernel_kind, int32_t mean_kind) {
  b200bo_handle_t h = nullptr;
  if (!out) return fail(h, B200BO_ERR_ARG, "null handle pointer");
  *out = nullptr;
  if (D < 1 || D > 32) return fail(h, B200BO_ERR_ARG, "input dimension D must be in 1..32");
  if (kernel_kind < 0 || kernel_kind > B200BO_KERNEL_MAT52ARD) return fail(h, B200BO_ERR_ARG, "unknown kernel kind");
  if (mean_kind != B200BO_MEAN_ZERO && mean_kind != B200BO_MEAN_CONST) return fail(h, B200BO_ERR_ARG, "unknown mean kind");
  if (capacity < 0) return fail(h, B200BO_ERR_ARG, "negative capacity");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(h, B200BO_ERR_CUDA, "no CUDA device: libb200bo has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(h, B200BO_ERR_ARG, "device index out of range");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(h, B200BO_ERR_CUDA, "cudaGetDeviceProperties failed");
  if (prop.major != 10) return fail(h, B200BO_ERR_CUDA, "libb200bo is built for sm_100a (B200) only");
  h = new (std::nothrow) b200bo_handle_s();
  if (!h) return fail(nullptr, B200BO_ERR_ALLOC, "out of host memory");
  h->device = device; h->D = D; h->kernel_kind = kernel_kind; h->mean_kind = mean_kind;
  h->fam = fam_of(kernel_kind); h->iso = iso_of(kernel_kind);
  h->num_sms = prop.multiProcessorCount;
  h->hp.ll.assign(h->iso ? 1 : D, 0.0);
  cudaSetDevice(device);
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  if (cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
      cudaStreamCreateWithPriority(&h->stream2, cudaStreamNonBlocking, prio_lo) != cudaSuccess) {
    delete h;
    return fail(nullptr, B200BO_ERR_CUDA, "stream create failed");
  }
  for (auto& e : h->ev) cudaEventCreate(&e);
  int32_t rc = alloc_device(h, capacity);
  if (rc != B200BO_OK) {
    g_err = h->err;
    free_device(h);
    for (auto& e : h->ev) if (e) cudaEventDestroy(e);
    cudaStreamDestroy(h->stream2);
    cudaStreamDestroy(h->stream);
    delete h;
    return rc;
  }
  h->fitted = true;   // empty model: prior
  *out = h;
  return B200BO_OK;
}

B200BO_API int32_t b200bo_destroy(b200bo_handle_t h) {
  if (!h) return B200BO_OK;
  for (auto* r : h->replicas) b200bo_destroy(r);
  h->replicas.clear();
  for (auto* wk : h->workers) b200bo_destroy(wk);
  h->workers.clear();
  if (h->joint) { b200bo_destroy(h->joint); h->joint = nullptr; }
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  nccl_destroy(h);
  cudaFree(h->drec); h->drec = nullptr;
  free_device(h);
  if (h->dio) cudaFree(h->dio);
  if (h->dlbub) cudaFree(h->dlbub);
  for (auto& e : h->ev) if (e) cudaEventDestroy(e);
  for (auto& e : h->syrk_ev) cudaEventDestroy(e);
  for (auto& e : h->la_ev) cudaEventDestroy(e);
  for (auto& e : h->fw_ev) cudaEventDestroy(e);
  for (auto& e : h->ch_ev) cudaEventDestroy(e);
  for (cudaStream_t st : {h->stream3, h->stream4, h->stream5}) if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
  for (auto& e : h->acq_ev) if (e) cudaEventDestroy(e);
  for (auto& e : h->gemm_ev) cudaEventDestroy(e);
  if (h->stream2) { cudaStreamSynchronize(h->stream2); cudaStreamDestroy(h->stream2); }
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return B200BO_OK;
}

B200BO_API int32_t b200bo_create_multi(b200bo_handle_t* out, int32_t n_gpus, const int32_t* devices, int32_t D, int64_t capacity,
                                       int32_t kernel_kind, int32_t mean_kind) {
  if (!out) return fail(nullptr, B200BO_ERR_ARG, "null handle pointer");
  *out = nullptr;
  if (n_gpus < 1 || n_gpus > 64) return fail(nullptr, B200BO_ERR_ARG, "n_gpus must be in 1..64");
  b200bo_handle_t h = nullptr;
  int32_t rc = b200bo_create(&h, devices ? devices[0] : 0, D, capacity, kernel_kind, mean_kind);
  if (rc != B200BO_OK) return rc;
  for (int r = 1; r < n_gpus; ++r) {
    b200bo_handle_t c = nullptr;
    rc = b200bo_create(&c, devices ? devices[r] : r, D, capacity, kernel_kind, mean_kind);
    if (rc != B200BO_OK) { b200bo_destroy(h); return rc; }
    c->is_replica = true;
    h->replicas.push_back(c);
  }
  if (n_gpus > 1) {
    std::string err;
    if (nccl_init_all(all_replicas(h), &err) != 0) { b200bo_destroy(h); return fail(nullptr, B200BO_ERR_NCCL, err); }
    for (auto* r : all_replicas(h)) {
      cudaSetDevice(r->device);
      if (comm_buffers(r, n_gpus) != cudaSuccess) { b200bo_destroy(h); return fail(nullptr, B200BO_ERR_CUDA, "exchange buffers"); }
    }
    cudaSetDevice(h->device);
  }
  *out = h;
  return B200BO_OK;
}

B200BO_API int32_t b200bo_num_gpus(b200bo_handle_t h, int32_t* n) {
  if (!h || !n) return fail(h, B200BO_ERR_ARG, "null argument");
  *n = h->replicas.empty() ? h->comm_world : 1 + (int32_t)h->replicas.size();
  return B200BO_OK;
}

B200BO_API int32_t b200bo_comm_unique_id(uint8_t* id128) {
  if (!id128) return fail(nullptr, B200BO_ERR_ARG, "null argument");
  std::string err;
  if (nccl_unique_id(id128, &err) != 0) return fail(nullptr, B200BO_ERR_NCCL, err);
  return B200BO_OK;
}

B200BO_API int32_t b200bo_comm_init_rank(b200bo_handle_t h, int32_t world, int32_t rank, const uint8_t* id128) {
  if (!h || !id128 || world < 1 || rank < 0 || rank >= world) return fail(h, B200BO_ERR_ARG, "bad arguments to comm_init_rank");
  if (!h->replicas.empty()) return fail(h, B200BO_ERR_STATE, "a multi handle already owns its communicators");
  if (h->comm) return fail(h, B200BO_ERR_STATE, "a communicator is already attached");
  cudaSetDevice(h->device);
  std::string err;
  if (nccl_init_rank(h, world, rank, id128, &err) != 0) return fail(h, B200BO_ERR_NCCL, err);
  CU(comm_buffers(h, world));
  return B200BO_OK;
}

B200BO_API int32_t b200bo_comm_destroy(b200bo_handle_t h) {
  if (!h) return fail(h, B200BO_ERR_ARG, "null handle");
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  nccl_destroy(h);
  return B200BO_OK;
}

B200BO_API int32_t b200bo_set_stream(b200bo_handle_t h, void* s) {
  if (!h) return fail(h, B200BO_ERR_ARG, "null handle");
  CU(cudaStreamSynchronize(h->stream));
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  if (s) { h->stream = (cudaStream_t)s; h->own_stream = false; }
  else { CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)); h->own_stream = true; }
  return B200BO_OK;
}

B200BO_API int32_t b200bo_sync(b200bo_handle_t h) {
  if (!h) return fail(h, B200BO_ERR_ARG, "null handle");
  CU(cudaStreamSynchronize(h->stream));
  return B200BO_OK;
}

B200BO_API int32_t b200bo_num_params(b200bo_handle_t h, int32_t* P) {
  if (!h || !P) return fail(h, B200BO_ERR_ARG, "null argument");
  *P = num_params(h);
  return B200BO_OK;
}

B200BO_API int32_t b200bo_set_params(b200bo_handle_t h, const double* th, int32_t P) {
  if (!h || !th) return fail(h, B200BO_ERR_ARG, "null argument");
  for (auto* r : h->replicas) { const int32_t rc = b200bo_set_params(r, th, P); if (rc) return fail(h, rc, r->err); }
  if (P != num_params(h)) return fail(h, B200BO_ERR_ARG, "parameter vector has the wrong length");
  for (int i = 0; i < P; ++i) if (!isfinite(th[i])) return fail(h, B200BO_ERR_ARG, "non-finite hyper-parameter");
  int i = 0;
  h->hp.lognoise = th[i++];
  if (h->mean_kind == B200BO_MEAN_CONST) h->hp.beta = th[i++];
  for (auto& l : h->hp.ll) l = th[i++];
  h->hp.lsigma = th[i++];
  h->fitted = (h->N == 0);
  cudaSetDevice(h->device);
  return sync_inv_ell(h);
}

// EXT GaussianProcesses.jl: "non-flat priors can be specified directly on the GP parameters" (reference src/models/gp.jl:30-35,
// set_priors!(obj, [Normal(mu, sigma), ...])); the MAP target is then mll + sum of the log prior densities and dtarget gains their
// derivatives.  kind[i] = 0 flat (default), 1 Normal(a[i], b[i]), in the full parameter order [logNoise, (beta), kernel...].
B200BO_API int32_t b200bo_set_priors(b200bo_handle_t h, int32_t P, const int32_t* kind, const double* a, const double* b) {
  if (!h) return fail(h, B200BO_ERR_ARG, "null handle");
  if (P == 0) { h->prior_kind.clear(); h->prior_a.clear(); h->prior_b.clear(); }
  else {
    if (P != num_params(h) || !kind || !a || !b) return fail(h, B200BO_ERR_ARG, "priors need one entry per parameter");
    for (int i = 0; i < P; ++i) {
      if (kind[i] != 0 && kind[i] != 1) return fail(h, B200BO_ERR_ARG, "unknown prior kind (0 = flat, 1 = Normal)");
      if (kind[i] == 1 && !(b[i] > 0.0)) return fail(h, B200BO_ERR_ARG, "Normal prior needs sigma > 0");
    }
    h->prior_kind.assign(kind, kind + P); h->prior_a.assign(a, a + P); h->prior_b.assign(b, b + P);
  }
  for (auto* r : h->replicas) b200bo_set_priors(r, P, kind, a, b);
  return B200BO_OK;
}

B200BO_API int32_t b200bo_get_params(b200bo_handle_t h, double* th, int32_t P) {
  if (!h || !th) return fail(h, B200BO_ERR_ARG, "null argument");
  if (P != num_params(h)) return fail(h, B200BO_ERR_ARG, "parameter vector has the wrong length");
  int i = 0;
  th[i++] = h->hp.lognoise;
  if (h->mean_kind == B200BO_MEAN_CONST) th[i++] = h->hp.beta;
  for (auto l : h->hp.ll) th[i++] = l;
  th[i++] = h->hp.lsigma;
  return B200BO_OK;
}

B200BO_API int32_t b200bo_fit(b200bo_handle_t h, const double* X, const double* y, int64_t N) {
  if (!h || N < 0 || (N > 0 && (!X || !y))) return fail(h, B200BO_ERR_ARG, "bad arguments to fit");
  if (!h->replicas.empty()) {     // every replica factors the same data: identical kernels on identical inputs, identical bits
    std::vector<b200bo_handle_s*> kids;
    kids.swap(h->replicas);
    std::vector<int32_t> rc(kids.size() + 1, B200BO_OK);
    std::vector<std::thread> th;
    for (size_t r = 0; r < kids.size(); ++r) th.emplace_back([&, r] { rc[r + 1] = b200bo_fit(kids[r], X, y, N); });
    rc[0] = b200bo_fit(h, X, y, N);
    for (auto& t : th) t.join();
    h->replicas.swap(kids);
    for (size_t r = 0; r < rc.size(); ++r) if (rc[r]) return r ? fail(h, rc[r], h->replicas[r - 1]->err) : rc[r];
    return B200BO_OK;
  }
  cudaSetDevice(h->device);
  h->hX.assign(X, X + N * h->D);
  h->hy.assign(y, y + N);
  h->N = N;
  h->data_version++;
  int32_t rc = upload_data(h);
  if (rc) return rc;
  return refit(h);
}

// elastic append of m points (EXT ElasticPDMats append!, reached from src/models/gp.jl:11; `repetitions` columns at
// src/BayesianOptimization.jl:194-196): ONE enqueue -- per point the new factor row (forward solve), pivot and block-inverse row; alpha
// and the log-determinant once behind the last point -- and ONE synchronisation.  *ok = false when positive definiteness was lost on
// the way (the caller refactors with the jitter rule).
static int32_t append_elastic(b200bo_handle_t h, const double* Xn, const double* yn, int64_t m, bool* ok) {
  const int64_t D = h->D, N0 = h->N;
  *ok = true;
  CU(cudaMemcpyAsync(h->dX + N0 * D, Xn, sizeof(double) * m * D, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->dy + N0, yn, sizeof(double) * m, cudaMemcpyHostToDevice, h->stream));
  CU(launch_scale_inputs(h, N0, N0 + m));
  CU(cudaMemsetAsync(h->dinfo, 0, sizeof(int), h->stream));
  int64_t j = 0;
  if (h->wt_valid && !h->lite) {
    // the acquisition path left W = L^-1 behind: the rows of all new points that fit the current 128-block come from ONE dense product
    // (append.cu: launch_append_block) instead of one chained forward solve per point
    const int64_t room = NB - (N0 % NB);                       // N0 % NB == 0: a fresh block
    const int mb = (int)std::min<int64_t>(m, room);
    CU(launch_append_block(h, h->noise_total, mb));            // advances h->N / h->Np
    j = mb;
    h->wt_valid = false;                                       // the factor grew
    if (j == m) {
      CU(launch_backward_solve(h, h->dz, h->dw, h->dalpha, (int)(h->Np / NB)));
      CU(launch_logdet_dot(h));
    }
  }
  for (; j < m; ++j) CU(launch_append_one(h, h->noise_total, j == m - 1));          // advances h->N / h->Np
  int info = 0;
  double sc[2];
  CU(cudaMemcpyAsync(&info, h->dinfo, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(sc, h->dscal, sizeof(sc), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if (info != 0) *ok = false;
  else h->mll = -0.5 * (sc[1] + sc[0] + (double)h->N * 1.8378770664093453);
  return B200BO_OK;
}

B200BO_API int32_t b200bo_append(b200bo_handle_t h, const double* Xn, const double* yn, int64_t m) {
  if (!h || m < 0 || (m > 0 && (!Xn || !yn))) return fail(h, B200BO_ERR_ARG, "bad arguments to append");
  if (m == 0) return B200BO_OK;
  if (!h->replicas.empty()) {
    std::vector<b200bo_handle_s*> kids;
    kids.swap(h->replicas);
    std::vector<int32_t> rc(kids.size() + 1, B200BO_OK);
    std::vector<std::thread> th;
    for (size_t r = 0; r < kids.size(); ++r) th.emplace_back([&, r] { rc[r + 1] = b200bo_append(kids[r], Xn, yn, m); });
    rc[0] = b200bo_append(h, Xn, yn, m);
    for (auto& t : th) t.join();
    h->replicas.swap(kids);
    for (size_t r = 0; r < rc.size(); ++r) if (rc[r]) return r ? fail(h, rc[r], h->replicas[r - 1]->err) : rc[r];
    return B200BO_OK;
  }
  cudaSetDevice(h->device);
  h->data_version++;
  const int64_t D = h->D, N0 = (int64_t)h->hy.size();
  // Elastic path (EXT ElasticPDMats append!): a valid factor, room in the buffers, a handful of new points.
  const bool elastic = h->fitted && h->N > 0 && m <= 16 && h->N + m <= h->cap;
  if (elastic) {
    bool ok = true;
    const int32_t rc = append_elastic(h, Xn, yn, m, &ok);
    if (rc != B200BO_OK) {
      // a CUDA error half-way: the host copy was not touched yet; forget the partial device state (h->N is the count every
      // query reports and sizes caller buffers from) and refactor the old data lazily
      h->N = N0;
      h->need_upload = true;
      h->fitted = false;
      h->acq_ready = 0; h->wt_valid = false;
      return rc;
    }
    if (ok) {
      h->hX.insert(h->hX.end(), Xn, Xn + m * D);
      h->hy.insert(h->hy.end(), yn, yn + m);
      h->acq_ready = 0; h->wt_valid = false;     // the factor grew: W = L^-1 and its slices are rebuilt by the next acquisition
      return B200BO_OK;
    }
  }
  h->hX.insert(h->hX.end(), Xn, Xn + m * D);
  h->hy.insert(h->hy.end(), yn, yn + m);
  h->N = (int64_t)h->hy.size();
  int32_t rc = upload_data(h);
  if (rc) return rc;
  return refit(h);
}

B200BO_API int32_t b200bo_refit(b200bo_handle_t h) {
  if (!h) return fail(h, B200BO_ERR_ARG, "null handle");
  return on_replicas(h, [](b200bo_handle_s* r, int) { cudaSetDevice(r->device); return refit(r); });
}

B200BO_API int32_t b200bo_dims(b200bo_handle_t h, int32_t* D, int64_t* N) {
  if (!h) return fail(h, B200BO_ERR_ARG, "null handle");
  if (D) *D = h->D;
  if (N) *N = h->N;
  return B200BO_OK;
}

B200BO_API int32_t b200bo_maxy(b200bo_handle_t h, double* m) {
  if (!h || !m) return fail(h, B200BO_ERR_ARG, "null argument");
  double v = -INFINITY;
  for (double y : h->hy) v = y > v ? y : v;
  *m = v;
  return B200BO_OK;
}

B200BO_API int32_t b200bo_get_data(b200bo_handle_t h, double* X, double* y) {
  if (!h) return fail(h, B200BO_ERR_ARG, "null handle");
  // exactly the N points b200bo_dims reports (the caller sizes its buffers from that)
  if (X && h->N > 0) memcpy(X, h->hX.data(), sizeof(double) * h->N * h->D);
  if (y && h->N > 0) memcpy(y, h->hy.data(), sizeof(double) * h->N);
  return B200BO_OK;
}

B200BO_API int32_t b200bo_get_mll(b200bo_handle_t h, double* mll) {
  if (!h || !mll) return fail(h, B200BO_ERR_ARG, "null argument");
  cudaSetDevice(h->device);
  int32_t rc = ensure_fitted(h);
  if (rc) return rc;
  *mll = h->mll;
  return B200BO_OK;
}

B200BO_API int32_t b200bo_get_alpha(b200bo_handle_t h, double* alpha) {
  if (!h || !alpha) return fail(h, B200BO_ERR_ARG, "null argument");
  cudaSetDevice(h->device);
  int32_t rc = ensure_fitted(h);
  if (rc) return rc;
  if (h->N) CU(cudaMemcpyAsync(alpha, h->dalpha, sizeof(double) * h->N, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return B200BO_OK;
}

B200BO_API int32_t b200bo_get_factor(b200bo_handle_t h, double* U) {
  if (!h || !U) return fail(h, B200BO_ERR_ARG, "null argument");
  cudaSetDevice(h->device);
  int32_t rc = ensure_fitted(h);
  if (rc) return rc;
  const int64_t N = h->N;
  if (N == 0) return B200BO_OK;
  // row-major L (lower) == column-major U (upper); zero the mirrored other half on the host
  CU(cudaMemcpy2DAsync(U, sizeof(double) * N, h->dL, sizeof(double) * h->ld, sizeof(double) * N, N, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  for (int64_t r = 0; r < N; ++r)           // memory row r = column r of U: entries below the diagonal are zero
    for (int64_t c = r + 1; c < N; ++c) U[r * N + c] = 0.0;
  return B200BO_OK;
}

B200BO_API int32_t b200bo_jitter_tries(b200bo_handle_t h, int32_t* t) {
  if (!h || !t) return fail(h, B200BO_ERR_ARG, "null argument");
  *t = h->jitter;
  return B200BO_OK;
}

B200BO_API int32_t b200bo_kmat_dev(b200bo_handle_t h, double* dK, int64_t ld) {
  if (!h || !dK) return fail(h, B200BO_ERR_ARG, "null argument");
  if (ld < h->N || (ld & 1) || ((uintptr_t)dK & 15)) return fail(h, B200BO_ERR_ARG, "dK must be 16-byte aligned with even ld >= N");
  cudaSetDevice(h->device);
  std::vector<double> ie;
  upload_inv_ell(h, ie);
  CU(cudaMemcpyAsync(h->dinv_ell, ie.data(), sizeof(double) * 2 * h->D, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));   // ie is a host temporary
  CU(launch_scale_inputs(h, 0, h->Np));
  const double noise = exp(2.0 * h->hp.lognoise) + std::numeric_limits<double>::epsilon();
  CU(cudaEventRecord(h->ev[0], h->stream));
  CU(launch_kmat(h, dK, ld, h->N, h->Np, noise, false));
  CU(cudaEventRecord(h->ev[1], h->stream));
  return B200BO_OK;
}

B200BO_API int32_t b200bo_kmat(b200bo_handle_t h, double* K) {
  if (!h || !K) return fail(h, B200BO_ERR_ARG, "null argument");
  const int64_t N = h->N;
  if (N == 0) return B200BO_OK;
  cudaSetDevice(h->device);
  const int64_t ldk = (N + 1) & ~(int64_t)1;
  int32_t rc = ensure_io(h, sizeof(double) * ldk * N);
  if (rc) return rc;
  rc = b200bo_kmat_dev(h, h->dio, ldk);
  if (rc) return rc;
  CU(cudaMemcpy2DAsync(K, sizeof(double) * N, h->dio, sizeof(double) * ldk, sizeof(double) * N, N, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  cudaEventElapsedTime(&h->timing[B200BO_T_KMAT], h->ev[0], h->ev[1]);
  return B200BO_OK;
}

// enqueue the acquisition step; with a communicator attached the local best is packed into this rank's exchange record and, when
// `exchange` is set (one process per GPU), gathered from all ranks and merged: *dbest then holds the GLOBAL best on every rank
struct AcqHostIO { const double* Xs = nullptr; double *values = nullptr, *grad = nullptr, *mu = nullptr, *var = nullptr; };

static int32_t acquire_dev_impl(b200bo_handle_t h, int32_t kind, const double* p, int32_t np, const double* dXs, int64_t M, uint64_t seed,
                                int64_t idx_offset, double* dvalues, double* dgrad, double* dmu, double* dvar, b200bo_best_t* dbest, bool exchange,
                                const AcqHostIO* io = nullptr) {
  if (!h || M < 0 || (M > 0 && !dXs)) return fail(h, B200BO_ERR_ARG, "bad arguments to acquire");
  int32_t rc = check_acq(h, kind, np, dgrad != nullptr);
  if (rc) return rc;
  if (np > 0 && !p) return fail(h, B200BO_ERR_ARG, "null acquisition parameters");
  cudaSetDevice(h->device);
  rc = ensure_fitted(h);
  if (rc) return rc;
  AcqLaunch l;
  l.acq_kind = kind; l.p0 = np > 0 ? p[0] : 0.0; l.p1 = np > 1 ? p[1] : 0.0; l.seed = seed; l.idx_offset = idx_offset;
  l.dXs = dXs; l.M = M; l.dvalues = dvalues; l.dgrad = dgrad; l.dmu = dmu; l.dvar = dvar; l.dbest = dbest;
  if (io) { l.hXs = io->Xs; l.hvalues = io->values; l.hgrad = io->grad; l.hmu = io->mu; l.hvar = io->var; }
  CU(cudaEventRecord(h->ev[4], h->stream));
  if (M == 0 && dbest) {
    const b200bo_best_t none = {-INFINITY, -1};
    CU(cudaMemcpyAsync(dbest, &none, sizeof(none), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
  }
  CU(launch_acquire(h, l));
  if (h->comm && dbest) {
    CU(comm_buffers(h, h->comm_world));
    CU(launch_pack_best(h, dbest, dXs, idx_offset));
    if (exchange) {
      std::string err;
      if (nccl_allgather_records(h, &err) != 0) return fail(h, B200BO_ERR_NCCL, err);
      double* merged = h->drec + (size_t)B200BO_REC_DOUBLES * (size_t)(1 + h->comm_world);
      CU(launch_merge_best(h, h->comm_world, reinterpret_cast<b200bo_best_t*>(merged), merged + 2));
      CU(cudaMemcpyAsync(dbest, merged, sizeof(b200bo_best_t), cudaMemcpyDeviceToDevice, h->stream));
    }
  }
  CU(cudaEventRecord(h->ev[5], h->stream));
  return B200BO_OK;
}

B200BO_API int32_t b200bo_acquire_dev(b200bo_handle_t h, int32_t kind, const double* p, int32_t np, const double* dXs, int64_t M, uint64_t seed,
                           int64_t idx_offset, double* dvalues, double* dgrad, double* dmu, double* dvar, b200bo_best_t* dbest) {
  const bool exchange = h && h->comm && h->replicas.empty() && !h->is_replica && !h->in_multi;
  return acquire_dev_impl(h, kind, p, np, dXs, M, seed, idx_offset, dvalues, dgrad, dmu, dvar, dbest, exchange);
}

B200BO_API int32_t b200bo_predict_dev(b200bo_handle_t h, const double* dXs, int64_t M, double* dmu, double* dvar) {
  if (!h || M < 0 || (M > 0 && !dXs)) return fail(h, B200BO_ERR_ARG, "bad arguments to predict");
  cudaSetDevice(h->device);
  int32_t rc = ensure_fitted(h);
  if (rc) return rc;
  AcqLaunch l;
  l.acq_kind = -1; l.dXs = dXs; l.M = M; l.dmu = dmu; l.dvar = dvar;
  CU(cudaEventRecord(h->ev[4], h->stream));
  CU(launch_acquire(h, l));
  CU(cudaEventRecord(h->ev[5], h->stream));
  return B200BO_OK;
}

// host-pointer acquisition on ONE device: stage the candidates, enqueue (no synchronisation); outputs are copied back asynchronously
static int32_t acquire_host_enqueue(b200bo_handle_t h, int32_t kind, const double* p, int32_t np, const double* Xs, int64_t M, uint64_t seed,
                                    int64_t idx_offset, double* values, double* grad, double* mu, double* var, bool exchange,
                                    b200bo_best_t** dbest_out) {
  cudaSetDevice(h->device);
  const int64_t D = h->D;
  // staging layout: Xs [M*D] | values [M] | mu [M] | var [M] | grad [M*D] | best
  const int64_t nd = M * D + 3 * M + (grad ? M * D : 0) + 2;
  int32_t rc = ensure_io(h, sizeof(double) * nd);
  if (rc) return rc;
  double* dXs = h->dio;
  double* dval = dXs + M * D;
  double* dmu = dval + M;
  double* dvar = dmu + M;
  double* dgrad = grad ? dvar + M : nullptr;
  b200bo_best_t* dbest = reinterpret_cast<b200bo_best_t*>(dvar + M + (grad ? M * D : 0));
  // With PINNED host buffers the transfers are issued by the acquisition step itself, chunk by chunk on its stream lanes (H2D of chunk
  // c+1 and D2H of chunk c-1 run under the kernels of chunk c).  Pageable buffers make cudaMemcpyAsync block the enqueuing thread, which
  // would stall the chunk pipeline: those are copied in one piece before / after the step.
  auto pinned = [](const void* ptr) {
    if (!ptr) return true;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
  };
  const bool pipelined = M > 0 && pinned(Xs) && pinned(values) && pinned(grad) && pinned(mu) && pinned(var);
  AcqHostIO io;
  io.Xs = Xs; io.values = values; io.grad = grad; io.mu = mu; io.var = var;
  if (M > 0 && !pipelined) CU(cudaMemcpyAsync(dXs, Xs, sizeof(double) * M * D, cudaMemcpyHostToDevice, h->stream));
  rc = acquire_dev_impl(h, kind, p, np, dXs, M, seed, idx_offset, dval, dgrad, mu ? dmu : nullptr, var ? dvar : nullptr, dbest, exchange, pipelined ? &io : nullptr);
  if (rc) return rc;
  if (M > 0 && !pipelined) {
    if (values) CU(cudaMemcpyAsync(values, dval, sizeof(double) * M, cudaMemcpyDeviceToHost, h->stream));
    if (mu) CU(cudaMemcpyAsync(mu, dmu, sizeof(double) * M, cudaMemcpyDeviceToHost, h->stream));
    if (var) CU(cudaMemcpyAsync(var, dvar, sizeof(double) * M, cudaMemcpyDeviceToHost, h->stream));
    if (grad) CU(cudaMemcpyAsync(grad, dgrad, sizeof(double) * M * D, cudaMemcpyDeviceToHost, h->stream));
  }
  *dbest_out = dbest;
  return B200BO_OK;
}

B200BO_API int32_t b200bo_acquire(b200bo_handle_t h, int32_t kind, const double* p, int32_t np, const double* Xs, int64_t M, uint64_t seed,
                       int64_t idx_offset, double* values, double* grad, double* mu, double* var, b200bo_best_t* best, double* best_x) {
  if (!h || M < 0 || (M > 0 && !Xs)) return fail(h, B200BO_ERR_ARG, "bad arguments to acquire");
  int32_t rc = check_acq(h, kind, np, grad != nullptr);
  if (rc) return rc;
  const int64_t D = h->D;
  b200bo_best_t hb = {-INFINITY, -1};
  if (!h->replicas.empty()) {
    // ---- one process, several GPUs: contiguous column blocks, one host thread per replica, ONE grouped all-gather, merge on device 0 ----
    const std::vector<b200bo_handle_s*> reps = all_replicas(h);
    const int R = (int)reps.size();
    std::vector<b200bo_best_t*> dbest(R, nullptr);
    rc = on_replicas(h, [&](b200bo_handle_s* r, int k) {
      int64_t lo, hi; shard_bounds(M, R, k, &lo, &hi);
      return acquire_host_enqueue(r, kind, p, np, Xs + lo * D, hi - lo, seed, idx_offset + lo, values ? values + lo : nullptr, grad ? grad + lo * D : nullptr,
                                  mu ? mu + lo : nullptr, var ? var + lo : nullptr, false, &dbest[k]);
    });
    if (rc) return rc;
    std::string err;
    if (nccl_allgather_group(reps, &err) != 0) return fail(h, B200BO_ERR_NCCL, err);
    cudaSetDevice(h->device);
    double* merged = h->drec + (size_t)B200BO_REC_DOUBLES * (size_t)(1 + R);
    CU(launch_merge_best(h, R, reinterpret_cast<b200bo_best_t*>(merged), merged + 2));
    double hm[B200BO_REC_DOUBLES];
    CU(cudaMemcpyAsync(hm, merged, sizeof(double) * (2 + D), cudaMemcpyDeviceToHost, h->stream));
    for (auto* r : reps) { cudaSetDevice(r->device); CU(cudaStreamSynchronize(r->stream)); }
    cudaSetDevice(h->device);
    cudaEventElapsedTime(&h->timing[B200BO_T_ACQ], h->ev[4], h->ev[5]);
    memcpy(&hb, hm, sizeof(hb));
    if (best) *best = hb;
    if (best_x && hb.index >= 0) memcpy(best_x, hm + 2, sizeof(double) * D);
    return B200BO_OK;
  }
  const bool exchange = h->comm && !h->is_replica && !h->in_multi;
  b200bo_best_t* dbest = nullptr;
  rc = acquire_host_enqueue(h, kind, p, np, Xs, M, seed, idx_offset, values, grad, mu, var, exchange, &dbest);
  if (rc) return rc;
  double hm[B200BO_REC_DOUBLES];
  if (exchange) {     // the global best and ITS point (the winner may live on another rank)
    CU(cudaMemcpyAsync(hm, h->drec + (size_t)B200BO_REC_DOUBLES * (size_t)(1 + h->comm_world), sizeof(double) * (2 + D), cudaMemcpyDeviceToHost, h->stream));
  } else {
    CU(cudaMemcpyAsync(&hb, dbest, sizeof(hb), cudaMemcpyDeviceToHost, h->stream));
  }
  CU(cudaStreamSynchronize(h->stream));
  cudaEventElapsedTime(&h->timing[B200BO_T_ACQ], h->ev[4], h->ev[5]);
  if (exchange) {
    memcpy(&hb, hm, sizeof(hb));
    if (best_x && hb.index >= 0) memcpy(best_x, hm + 2, sizeof(double) * D);
  } else if (best_x && hb.index >= 0) {
    memcpy(best_x, Xs + (hb.index - idx_offset) * D, sizeof(double) * D);
  }
  if (best) *best = hb;
  return B200BO_OK;
}

// host-side merge of per-replica bests (the sharded search helpers): largest value, then lowest global index
static void merge_host(b200bo_best_t& acc, double* acc_x, const b200bo_best_t& b, const double* bx, int D) {
  if (b.index >= 0 && b.value == b.value && (acc.index < 0 || b.value > acc.value || (b.value == acc.value && b.index < acc.index))) {
    acc = b;
    if (acc_x && bx) memcpy(acc_x, bx, sizeof(double) * D);
  }
}

static int32_t upload_bounds(b200bo_handle_t h, const double* lb, const double* ub) {
  if (!lb || !ub) return fail(h, B200BO_ERR_ARG, "null bounds");
  for (int d = 0; d < h->D; ++d)
    if (!(lb[d] <= ub[d])) return fail(h, B200BO_ERR_ARG, "mins[i] should not exceed maxs[i]");      // utils.jl:106-107
  if (!h->dlbub) CU(cudaMalloc(&h->dlbub, sizeof(double) * 2 * h->D));
  std::vector<double> b(lb, lb + h->D);
  b.insert(b.end(), ub, ub + h->D);
  CU(cudaMemcpyAsync(h->dlbub, b.data(), sizeof(double) * 2 * h->D, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return B200BO_OK;
}

B200BO_API int32_t b200bo_lhs(b200bo_handle_t h, const double* lb, const double* ub, int64_t n_total, int64_t offset, int64_t n_local,
                              uint64_t seed, double* Xs) {
  if (!h || n_total < 0 || offset < 0 || n_local < 0 || offset + n_local > n_total || (n_local > 0 && !Xs))
    return fail(h, B200BO_ERR_ARG, "bad arguments to lhs");
  cudaSetDevice(h->device);
  int32_t rc = upload_bounds(h, lb, ub);
  if (rc) return rc;
  rc = ensure_io(h, sizeof(double) * (n_local * h->D + 2));
  if (rc) return rc;
  CU(launch_lhs(h, h->dio, n_total, offset, n_local, seed, h->dlbub));
  if (n_local > 0) CU(cudaMemcpyAsync(Xs, h->dio, sizeof(double) * n_local * h->D, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return B200BO_OK;
}

B200BO_API int32_t b200bo_acquire_lhs(b200bo_handle_t h, int32_t kind, const double* p, int32_t np, const double* lb, const double* ub,
                                      int64_t n_total, int64_t offset, int64_t n_local, uint64_t lhs_seed, uint64_t ts_seed,
                                      double* values, b200bo_best_t* best, double* best_x) {
  if (!h || n_total < 0 || offset < 0 || n_local < 0 || offset + n_local > n_total) return fail(h, B200BO_ERR_ARG, "bad arguments to acquire_lhs");
  int32_t rc = check_acq(h, kind, np, false);
  if (rc) return rc;
  if (!h->replicas.empty()) {     // every replica generates and scores its own block of the ONE global design
    const int R = 1 + (int)h->replicas.size();
    std::vector<b200bo_best_t> bs(R, b200bo_best_t{-INFINITY, -1});
    std::vector<double> bx((size_t)R * h->D, 0.0);
    rc = on_replicas(h, [&](b200bo_handle_s* r, int k) {
      int64_t lo, hi; shard_bounds(n_local, R, k, &lo, &hi);
      SoloScope solo(r);
      const int32_t e = b200bo_acquire_lhs(r, kind, p, np, lb, ub, n_total, offset + lo, hi - lo, lhs_seed, ts_seed, values ? values + lo : nullptr, &bs[k],
                                           bx.data() + (size_t)k * r->D);
      return e;
    });
    if (rc) return rc;
    b200bo_best_t acc = {-INFINITY, -1};
    for (int k = 0; k < R; ++k) merge_host(acc, best_x, bs[k], bx.data() + (size_t)k * h->D, h->D);
    if (best) *best = acc;
    return B200BO_OK;
  }
  cudaSetDevice(h->device);
  rc = upload_bounds(h, lb, ub);
  if (rc) return rc;
  const int64_t D = h->D, M = n_local;
  rc = ensure_io(h, sizeof(double) * (M * D + M + 4));
  if (rc) return rc;
  double* dXs = h->dio;
  double* dval = dXs + M * D;
  b200bo_best_t* dbest = reinterpret_cast<b200bo_best_t*>(dval + M);
  CU(launch_lhs(h, dXs, n_total, offset, n_local, lhs_seed, h->dlbub));
  rc = acquire_dev_impl(h, kind, p, np, dXs, M, ts_seed, offset, values ? dval : nullptr, nullptr, nullptr, nullptr, dbest, false);   // the best of THIS block
  if (rc) return rc;
  b200bo_best_t hb = {-INFINITY, -1};
  if (values && M > 0) CU(cudaMemcpyAsync(values, dval, sizeof(double) * M, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(&hb, dbest, sizeof(hb), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  cudaEventElapsedTime(&h->timing[B200BO_T_ACQ], h->ev[4], h->ev[5]);
  if (best) *best = hb;
  if (best_x && hb.index >= 0) {
    CU(cudaMemcpyAsync(best_x, dXs + (hb.index - offset) * D, sizeof(double) * D, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
  }
  return B200BO_OK;
}

B200BO_API int32_t b200bo_acquire_ascent(b200bo_handle_t h, int32_t kind, const double* p, int32_t np, const double* Xs, int64_t M,
                                         const double* lb, const double* ub, int32_t steps, double step0, int64_t idx_offset,
                                         double* Xout, double* values, b200bo_best_t* best, double* best_x) {
  if (!h || M < 0 || (M > 0 && !Xs) || steps < 0 || !(step0 > 0.0)) return fail(h, B200BO_ERR_ARG, "bad arguments to acquire_ascent");
  int32_t rc = check_acq(h, kind, np, true);
  if (rc) return rc;
  if (!h->replicas.empty()) {
    const int R = 1 + (int)h->replicas.size();
    std::vector<b200bo_best_t> bs(R, b200bo_best_t{-INFINITY, -1});
    std::vector<double> bx((size_t)R * h->D, 0.0);
    rc = on_replicas(h, [&](b200bo_handle_s* r, int k) {
      int64_t lo, hi; shard_bounds(M, R, k, &lo, &hi);
      SoloScope solo(r);
      const int32_t e = b200bo_acquire_ascent(r, kind, p, np, Xs + lo * r->D, hi - lo, lb, ub, steps, step0, idx_offset + lo, Xout ? Xout + lo * r->D : nullptr,
                                              values ? values + lo : nullptr, &bs[k], bx.data() + (size_t)k * r->D);
      return e;
    });
    if (rc) return rc;
    b200bo_best_t acc = {-INFINITY, -1};
    for (int k = 0; k < R; ++k) merge_host(acc, best_x, bs[k], bx.data() + (size_t)k * h->D, h->D);
    if (best) *best = acc;
    return B200BO_OK;
  }
  cudaSetDevice(h->device);
  rc = upload_bounds(h, lb, ub);
  if (rc) return rc;
  rc = ensure_fitted(h);
  if (rc) return rc;
  const int64_t D = h->D;
  rc = ensure_io(h, sizeof(double) * (M * D * 4 + M * 4 + 4));
  if (rc) return rc;
  double* dX = h->dio;
  double* dwork = dX + M * D;                                       // val, grad, Xb, Gb, Fb, S
  double* dFb = dwork + M + 3 * M * D;
  b200bo_best_t* dbest = reinterpret_cast<b200bo_best_t*>(dwork + 3 * M + 3 * M * D);
  b200bo_best_t hb = {-INFINITY, -1};
  if (M > 0) {
    CU(cudaMemcpyAsync(dX, Xs, sizeof(double) * M * D, cudaMemcpyHostToDevice, h->stream));
    AcqLaunch l;
    l.acq_kind = kind; l.p0 = np > 0 ? p[0] : 0.0; l.p1 = np > 1 ? p[1] : 0.0; l.idx_offset = idx_offset; l.M = M; l.dbest = dbest;
    CU(cudaEventRecord(h->ev[4], h->stream));
    CU(launch_ascent(h, l, dX, dwork, h->dlbub, steps, step0));
    CU(cudaEventRecord(h->ev[5], h->stream));
    if (Xout) CU(cudaMemcpyAsync(Xout, dX, sizeof(double) * M * D, cudaMemcpyDeviceToHost, h->stream));
    if (values) CU(cudaMemcpyAsync(values, dFb, sizeof(double) * M, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(&hb, dbest, sizeof(hb), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    cudaEventElapsedTime(&h->timing[B200BO_T_ACQ], h->ev[4], h->ev[5]);
    if (best_x && hb.index >= 0) {
      CU(cudaMemcpyAsync(best_x, dX + (hb.index - idx_offset) * D, sizeof(double) * D, cudaMemcpyDeviceToHost, h->stream));
      CU(cudaStreamSynchronize(h->stream));
    }
  }
  if (best) *best = hb;
  return B200BO_OK;
}

B200BO_API int32_t b200bo_sobol(b200bo_handle_t h, const double* lb, const double* ub, uint64_t index0, int64_t n, double* Xs) {
  if (!h || n < 0 || (n > 0 && !Xs) || index0 + (uint64_t)n > (1ull << 32)) return fail(h, B200BO_ERR_ARG, "bad arguments to sobol (indices must stay below 2^32)");
  cudaSetDevice(h->device);
  int32_t rc = upload_bounds(h, lb, ub);
  if (rc) return rc;
  rc = ensure_io(h, sizeof(double) * (n * h->D + 2));
  if (rc) return rc;
  CU(launch_sobol(h, h->dio, index0, n, h->dlbub));
  if (n > 0) CU(cudaMemcpyAsync(Xs, h->dio, sizeof(double) * n * h->D, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return B200BO_OK;
}

B200BO_API int32_t b200bo_acquire_lbfgs(b200bo_handle_t h, int32_t kind, const double* p, int32_t np, const double* Xs, int64_t M, const double* lb,
                                        const double* ub, int32_t maxeval, double ftol_rel, double ftol_abs, double xtol_rel, double xtol_abs,
                                        double maxtime, double step0, int64_t idx_offset, double* Xout, double* values, double* evals,
                                        b200bo_best_t* best, double* best_x) {
  if (!h || M < 0 || (M > 0 && !Xs) || !(step0 > 0.0) || ftol_rel < 0 || ftol_abs < 0 || xtol_rel < 0 || xtol_abs < 0)
    return fail(h, B200BO_ERR_ARG, "bad arguments to acquire_lbfgs");
  int32_t rc = check_acq(h, kind, np, true);
  if (rc) return rc;
  if (!h->replicas.empty()) {
    const int R = 1 + (int)h->replicas.size();
    std::vector<b200bo_best_t> bs(R, b200bo_best_t{-INFINITY, -1});
    std::vector<double> bx((size_t)R * h->D, 0.0);
    rc = on_replicas(h, [&](b200bo_handle_s* r, int k) {
      int64_t lo, hi; shard_bounds(M, R, k, &lo, &hi);
      SoloScope solo(r);
      return b200bo_acquire_lbfgs(r, kind, p, np, Xs + lo * r->D, hi - lo, lb, ub, maxeval, ftol_rel, ftol_abs, xtol_rel, xtol_abs, maxtime, step0,
                                  idx_offset + lo, Xout ? Xout + lo * r->D : nullptr, values ? values + lo : nullptr, evals ? evals + lo : nullptr, &bs[k],
                                  bx.data() + (size_t)k * r->D);
    });
    if (rc) return rc;
    b200bo_best_t acc = {-INFINITY, -1};
    for (int k = 0; k < R; ++k) merge_host(acc, best_x, bs[k], bx.data() + (size_t)k * h->D, h->D);
    if (best) *best = acc;
    return B200BO_OK;
  }
  cudaSetDevice(h->device);
  rc = upload_bounds(h, lb, ub);
  if (rc) return rc;
  rc = ensure_fitted(h);
  if (rc) return rc;
  const int64_t D = h->D, B = lbfgs_state_doubles((int)D);
  rc = ensure_io(h, sizeof(double) * (M * D + M * B + M + M * D + M + 4));
  if (rc) return rc;
  double* dXe = h->dio;
  double* dwork = dXe + M * D;                                       // state | val | grad | evals
  double* dval = dwork + M * B;
  double* devals = dval + M + M * D;
  b200bo_best_t* dbest = reinterpret_cast<b200bo_best_t*>(devals + M);
  b200bo_best_t hb = {-INFINITY, -1};
  if (M > 0) {
    CU(cudaMemcpyAsync(dXe, Xs, sizeof(double) * M * D, cudaMemcpyHostToDevice, h->stream));
    AcqLaunch l;
    l.acq_kind = kind; l.p0 = np > 0 ? p[0] : 0.0; l.p1 = np > 1 ? p[1] : 0.0; l.idx_offset = idx_offset; l.M = M; l.dbest = dbest;
    LbfgsOpts o;
    o.maxeval = maxeval; o.ftol_rel = ftol_rel; o.ftol_abs = ftol_abs; o.xtol_rel = xtol_rel; o.xtol_abs = xtol_abs; o.step0 = step0;
    CU(cudaEventRecord(h->ev[4], h->stream));
    int rounds = 0;
    CU(launch_lbfgs(h, l, dXe, dwork, h->dlbub, o, maxtime, &rounds));
    CU(cudaEventRecord(h->ev[5], h->stream));
    if (Xout) CU(cudaMemcpyAsync(Xout, dXe, sizeof(double) * M * D, cudaMemcpyDeviceToHost, h->stream));
    if (values) CU(cudaMemcpyAsync(values, dval, sizeof(double) * M, cudaMemcpyDeviceToHost, h->stream));
    if (evals) CU(cudaMemcpyAsync(evals, devals, sizeof(double) * M, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(&hb, dbest, sizeof(hb), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    cudaEventElapsedTime(&h->timing[B200BO_T_ACQ], h->ev[4], h->ev[5]);
    if (best_x && hb.index >= 0) {
      CU(cudaMemcpyAsync(best_x, dXe + (hb.index - idx_offset) * D, sizeof(double) * D, cudaMemcpyDeviceToHost, h->stream));
      CU(cudaStreamSynchronize(h->stream));
    }
  }
  if (best) *best = hb;
  return B200BO_OK;
}

// The derivative-free global search of the reference (NLopt :GN_DIRECT_L, the default for ThompsonSamplingSimple -- src/acquisition.jl:7-9;
// reached through the derivative-free wrapper, :31-36).  The rectangle bookkeeping is host logic (direct.h); every iteration's new centres
// -- all potentially optimal rectangles, all their longest sides -- are scored by ONE fused acquisition launch.  Evaluation e draws its
// Thompson sample from the Philox stream at global index e (a fresh epsilon per closure call, as in the reference).  Runs on the handle's
// first GPU only (the batches are tens of points); with a communicator attached every rank runs the same deterministic search.
B200BO_API int32_t b200bo_acquire_direct(b200bo_handle_t h, int32_t kind, const double* p, int32_t np, const double* lb, const double* ub,
                                         int32_t maxeval, double maxtime, int32_t width, int32_t variant, uint64_t seed, double* Xtrace, double* ftrace,
                                         int32_t* evals_out, int32_t* batches_out, b200bo_best_t* best, double* best_x) {
  if (!h || !lb || !ub || (maxeval <= 0 && !(maxtime > 0.0)) || ((Xtrace || ftrace) && maxeval <= 0))
    return fail(h, B200BO_ERR_ARG, "bad arguments to acquire_direct (a search needs maxeval > 0 or maxtime > 0; a trace needs maxeval)");
  int32_t rc = check_acq(h, kind, np, false);
  if (rc) return rc;
  const int D = h->D;
  if (D > 64) return fail(h, B200BO_ERR_ARG, "acquire_direct supports D <= 64");
  if (variant != B200BO_DIRECT_L && variant != B200BO_DIRECT_ORIG) return fail(h, B200BO_ERR_ARG, "unknown DIRECT variant");
  for (int d = 0; d < D; ++d)
    if (!(lb[d] <= ub[d]) || !std::isfinite(lb[d]) || !std::isfinite(ub[d])) return fail(h, B200BO_ERR_ARG, "DIRECT needs finite bounds with lb <= ub");
  SoloScope solo(h);
  DirectL s;
  s.init(D, maxeval, width, variant);
  std::vector<double> pts, xs, vals;
  const auto t0 = std::chrono::steady_clock::now();
  int batches = 0;
  for (;;) {
    const int64_t n = s.ask(pts);
    if (n == 0) break;
    xs.resize((size_t)n * D); vals.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i)
      for (int d = 0; d < D; ++d) xs[(size_t)i * D + d] = lb[d] + pts[(size_t)i * D + d] * (ub[d] - lb[d]);
    b200bo_best_t* dbest = nullptr;
    rc = acquire_host_enqueue(h, kind, p, np, xs.data(), n, seed, s.evals, vals.data(), nullptr, nullptr, nullptr, false, &dbest);
    if (rc) return rc;
    CU(cudaStreamSynchronize(h->stream));
    if (Xtrace) memcpy(Xtrace + (size_t)s.evals * D, xs.data(), sizeof(double) * n * D);
    if (ftrace) memcpy(ftrace + s.evals, vals.data(), sizeof(double) * n);
    s.tell(pts, vals.data());
    ++batches;
    if (maxtime > 0.0 && std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() >= maxtime) break;
  }
  if (evals_out) *evals_out = (int32_t)s.evals;
  if (batches_out) *batches_out = batches;
  if (best) { best->value = s.best_f; best->index = s.best_eval; }
  if (best_x && s.best_f > -INFINITY)
    for (int d = 0; d < D; ++d) best_x[d] = lb[d] + s.best_c[d] * (ub[d] - lb[d]);
  return B200BO_OK;
}

// optimizemodel!(::MAPGPOptimizer, model) (src/models/gp.jl:54-77): box-bounded L-BFGS ascent of the marginal log-likelihood over the
// masked parameters, every evaluation = mll + dmll on the device (K1-K5 + K7).  R starts run in lock-step: one b200bo_mll_sweep of R
// settings per round (on a multi handle the settings shard over the GPUs).  The model is left at the best parameters found.
B200BO_API int32_t b200bo_map_fit(b200bo_handle_t h, const double* Theta0, int32_t P, int32_t R, int32_t mask, const double* lb, const double* ub,
                                  int32_t maxeval, double ftol_rel, double ftol_abs, double xtol_rel, double xtol_abs, double maxtime,
                                  double* theta_best, double* mll_best, int32_t* evals_out, int32_t* status_out) {
  if (!h || !Theta0 || !lb || !ub || R < 1 || P < 1 || !theta_best) return fail(h, B200BO_ERR_ARG, "bad arguments to map_fit");
  for (int k = 0; k < P; ++k) if (!(lb[k] <= ub[k])) return fail(h, B200BO_ERR_ARG, "lower bound above upper bound");
  const int64_t B = lbfgs_state_doubles(P);
  std::vector<double> state((size_t)R * B, 0.0), xe((size_t)R * P), f(R), g((size_t)R * P);
  for (int r = 0; r < R; ++r)
    for (int k = 0; k < P; ++k) xe[(size_t)r * P + k] = lb_clamp(Theta0[(size_t)r * P + k], lb[k], ub[k]);
  LbfgsOpts o;
  o.maxeval = maxeval; o.ftol_rel = ftol_rel; o.ftol_abs = ftol_abs; o.xtol_rel = xtol_rel; o.xtol_abs = xtol_abs; o.step0 = 0.02;
  const auto t0 = std::chrono::steady_clock::now();
  const int cap = maxeval > 0 ? maxeval : 100000;
  int rounds = 0;
  for (; rounds < cap; ++rounds) {
    int32_t rc = b200bo_mll_sweep(h, xe.data(), P, R, mask, f.data(), g.data());
    if (rc == B200BO_ERR_NOTPD) {      // the reference's closure throws inside NLopt (FORCED_STOP): treat the point as infeasible
      for (int r = 0; r < R; ++r) f[r] = -INFINITY;
      std::fill(g.begin(), g.end(), 0.0);
    } else if (rc) return rc;
    int running = 0;
    for (int r = 0; r < R; ++r) {
      double* st = state.data() + (size_t)r * B;
      lbfgs_step(st, xe.data() + (size_t)r * P, f[r], g.data() + (size_t)r * P, lb, ub, P, o);
      if (LbfgsState(st, P).status() == 0.0) ++running;
    }
    if (running == 0) { ++rounds; break; }
    if (maxtime > 0.0 && std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() >= maxtime) { ++rounds; break; }
  }
  int bestr = 0;
  for (int r = 1; r < R; ++r)
    if (LbfgsState(state.data() + (size_t)r * B, P).f() > LbfgsState(state.data() + (size_t)bestr * B, P).f()) bestr = r;
  LbfgsState sb(state.data() + (size_t)bestr * B, P);
  memcpy(theta_best, sb.x, sizeof(double) * P);
  if (mll_best) *mll_best = sb.f();
  if (evals_out) *evals_out = (int32_t)sb.evals();
  if (status_out) *status_out = (int32_t)sb.status();
  // leave the model at the optimum (masked parameters only)
  const bool m_noise = mask & B200BO_MASK_NOISE, m_mean = (mask & B200BO_MASK_MEAN) && h->mean_kind == B200BO_MEAN_CONST, m_kern = mask & B200BO_MASK_KERN;
  const int np_full = num_params(h);
  std::vector<double> th(np_full);
  int32_t rc = b200bo_get_params(h, th.data(), np_full);
  if (rc) return rc;
  int i = 0, j = 0;
  if (m_noise) th[i] = sb.x[j++];
  ++i;
  if (h->mean_kind == B200BO_MEAN_CONST) { if (m_mean) th[i] = sb.x[j++]; ++i; }
  if (m_kern) for (; i < np_full; ++i) th[i] = sb.x[j++];
  if (j != P) return fail(h, B200BO_ERR_ARG, "Theta0 has the wrong number of rows for this mask");
  return b200bo_set_params(h, th.data(), np_full);
}

B200BO_API int32_t b200bo_predict(b200bo_handle_t h, const double* Xs, int64_t M, double* mu, double* var) {
  if (!h || M < 0 || (M > 0 && (!Xs || !mu || !var))) return fail(h, B200BO_ERR_ARG, "bad arguments to predict");
  if (!h->replicas.empty()) {
    const int R = 1 + (int)h->replicas.size();
    return on_replicas(h, [&](b200bo_handle_s* r, int k) {
      int64_t lo, hi; shard_bounds(M, R, k, &lo, &hi);
      SoloScope solo(r);
      const int32_t e = b200bo_predict(r, Xs + lo * r->D, hi - lo, mu + lo, var + lo);
      return e;
    });
  }
  cudaSetDevice(h->device);
  const int64_t D = h->D;
  int32_t rc = ensure_io(h, sizeof(double) * (M * D + 2 * M + 2));
  if (rc) return rc;
  double* dXs = h->dio;
  double* dmu = dXs + M * D;
  double* dvar = dmu + M;
  if (M > 0) CU(cudaMemcpyAsync(dXs, Xs, sizeof(double) * M * D, cudaMemcpyHostToDevice, h->stream));
  rc = b200bo_predict_dev(h, dXs, M, dmu, dvar);
  if (rc) return rc;
  if (M > 0) {
    CU(cudaMemcpyAsync(mu, dmu, sizeof(double) * M, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(var, dvar, sizeof(double) * M, cudaMemcpyDeviceToHost, h->stream));
  }
  CU(cudaStreamSynchronize(h->stream));
  cudaEventElapsedTime(&h->timing[B200BO_T_ACQ], h->ev[4], h->ev[5]);
  return B200BO_OK;
}

// myrand(model, X::Matrix) (src/models/gp.jl:7): ONE joint draw from the posterior at the M columns of Xs -- EXT rand(gp, X) =
// mu + chol(make_posdef!(Sigma_post)) eps.  See joint.cu: the augmented covariance is factorised on a worker model by the fit kernels and
// its trailing block applied to the Philox normals eps_j = stream(seed, idx_offset + j).  tries = make_posdef! retries that were needed.
B200BO_API int32_t b200bo_rand_joint(b200bo_handle_t h, const double* Xs, int64_t M, uint64_t seed, int64_t idx_offset, double* sample,
                                     double* mu_out, int32_t* tries_out) {
  if (!h || M < 0 || (M > 0 && (!Xs || !sample))) return fail(h, B200BO_ERR_ARG, "bad arguments to rand_joint");
  if (tries_out) *tries_out = 0;
  if (M == 0) return B200BO_OK;
  const int64_t D = h->D, N = h->N;
  if (N + M > 32768) return fail(h, B200BO_ERR_ARG, "rand_joint: observations + sample points must not exceed 32768 (the joint covariance is dense)");
  std::vector<double> mu(M), var(M);
  int32_t rc;
  {
    SoloScope solo(h);                   // a dense M x M factor does not shard: the primary GPU draws the sample
    rc = b200bo_predict(h, Xs, M, mu.data(), var.data());
  }
  if (rc) return rc;
  cudaSetDevice(h->device);
  if (!h->joint) { rc = new_worker(h, std::max<int64_t>(N + M, NB), &h->joint); if (rc) return rc; }
  b200bo_handle_s* wk = h->joint;
  wk->hp = h->hp; wk->syrk_engine = h->syrk_engine;
  wk->hX.assign(h->hX.begin(), h->hX.begin() + N * D);
  wk->hX.insert(wk->hX.end(), Xs, Xs + M * D);
  wk->hy.assign(h->hy.begin(), h->hy.begin() + N);
  wk->hy.resize(N + M, h->mean_kind == B200BO_MEAN_CONST ? h->hp.beta : 0.0);
  wk->N = N + M;
  rc = upload_data(wk);
  if (rc) { h->err = wk->err; return rc; }
  {
    b200bo_handle_s* h = wk;             // CU() reports on the worker
    std::vector<double> ie;
    upload_inv_ell(h, ie);
    CU(cudaMemcpyAsync(h->dinv_ell, ie.data(), sizeof(double) * 2 * D, cudaMemcpyHostToDevice, h->stream));
    CU(launch_scale_inputs(h, 0, h->Np));
  }
  const double sf2 = exp(2.0 * h->hp.lsigma);
  const double noise = N > 0 && h->fitted ? h->noise_total : exp(2.0 * h->hp.lognoise) + std::numeric_limits<double>::epsilon();
  double mean_var = 0.0;
  for (int64_t i = 0; i < M; ++i) mean_var += var[i];
  mean_var /= (double)M;
  double jit = 0.0;
  int tries = 0;
  for (;;) {
    b200bo_handle_s* h = wk;
    CU(launch_kmat(h, h->dL, h->ld, h->N, h->Np, noise, true));
    CU(launch_joint_diag(h, N, M, sf2 + jit));
    CU(launch_cholesky(h));
    int info = 0;
    CU(cudaMemcpyAsync(&info, h->dinfo, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (info == 0) break;
    if (tries >= 10) return fail(h, B200BO_ERR_NOTPD, "posterior covariance not positive definite after 10 jitter retries");
    jit += 1e-6 * (mean_var + jit);     // make_posdef!: 1e-6 tr(Sigma_post)/M, the trace including the jitter added so far
    ++tries;
  }
  {
    b200bo_handle_s* h = wk;
    // dw: mu | dz: eps | dalpha: sample   (the worker's single-RHS vectors; capacity >= N + M)
    CU(cudaMemcpyAsync(h->dw, mu.data(), sizeof(double) * M, cudaMemcpyHostToDevice, h->stream));
    CU(launch_joint_sample(h, N, M, h->dw, h->dz, h->dalpha, seed, idx_offset));
    CU(cudaMemcpyAsync(sample, h->dalpha, sizeof(double) * M, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
  }
  h->launches += wk->launches; wk->launches = 0;
  if (mu_out) memcpy(mu_out, mu.data(), sizeof(double) * M);
  if (tries_out) *tries_out = tries;
  return B200BO_OK;
}

static int32_t mll_sweep_impl(b200bo_handle_t h, const double* Theta, int32_t P, int32_t S, int32_t mask, double* mll, double* dmll);

B200BO_API int32_t b200bo_mll_sweep(b200bo_handle_t h, const double* Theta, int32_t P, int32_t S, int32_t mask, double* mll, double* dmll) {
  const int32_t rc = mll_sweep_impl(h, Theta, P, S, mask, mll, dmll);
  if (rc != B200BO_OK || h->prior_kind.empty() || h->lite || h->is_replica || h->in_multi) return rc;
  // target = mll + log prior of the swept parameters (the priors of the parameters held fixed are constants of the sweep)
  const bool m_noise = mask & B200BO_MASK_NOISE, m_mean = (mask & B200BO_MASK_MEAN) && h->mean_kind == B200BO_MEAN_CONST, m_kern = mask & B200BO_MASK_KERN;
  std::vector<int> full;                                      // swept row -> index in the full parameter vector
  int i = 0;
  if (m_noise) full.push_back(i);
  ++i;
  if (h->mean_kind == B200BO_MEAN_CONST) { if (m_mean) full.push_back(i); ++i; }
  if (m_kern) for (; i < num_params(h); ++i) full.push_back(i);
  for (int s = 0; s < S; ++s)
    for (int r = 0; r < P; ++r) {
      const int k = full[r];
      if (h->prior_kind[k] != 1) continue;
      const double z = (Theta[(int64_t)s * P + r] - h->prior_a[k]) / h->prior_b[k];
      mll[s] += -0.5 * z * z - log(h->prior_b[k]) - 0.9189385332046727;
      if (dmll) dmll[(int64_t)s * P + r] += -z / h->prior_b[k];
    }
  return B200BO_OK;
}

static int32_t mll_sweep_impl(b200bo_handle_t h, const double* Theta, int32_t P, int32_t S, int32_t mask, double* mll, double* dmll) {
  if (!h || !Theta || !mll || S < 0) return fail(h, B200BO_ERR_ARG, "bad arguments to mll_sweep");
  const bool m_noise = mask & B200BO_MASK_NOISE, m_mean = (mask & B200BO_MASK_MEAN) && h->mean_kind == B200BO_MEAN_CONST,
             m_kern = mask & B200BO_MASK_KERN;
  const int nl = (int)h->hp.ll.size();
  const int Pexp = (m_noise ? 1 : 0) + (m_mean ? 1 : 0) + (m_kern ? nl + 1 : 0);
  if (P != Pexp) return fail(h, B200BO_ERR_ARG, "Theta has the wrong number of rows for this mask");
  if (h->N == 0) return fail(h, B200BO_ERR_STATE, "mll needs observations");
  if (!h->replicas.empty()) {     // the settings shard over the replicas in contiguous blocks
    const int R = 1 + (int)h->replicas.size();
    return on_replicas(h, [&](b200bo_handle_s* r, int k) {
      int64_t lo, hi; shard_bounds(S, R, k, &lo, &hi);
      SoloScope solo(r);
      const int32_t e = hi > lo ? b200bo_mll_sweep(r, Theta + lo * P, P, (int32_t)(hi - lo), mask, mll + lo, dmll ? dmll + lo * P : nullptr) : B200BO_OK;
      return e;
    });
  }
  cudaSetDevice(h->device);
  if (!h->lite && h->sweep_workers > 0) {
    // ---- several settings in flight: K worker models on this device, contiguous blocks of settings, one host thread each ----
    const int K = std::max(1, std::min<int>(h->sweep_workers, S));
    while ((int)h->workers.size() < K) {
      b200bo_handle_s* wk = nullptr;
      const int32_t rcw = new_worker(h, h->cap, &wk);
      if (rcw != B200BO_OK) return rcw;
      wk->fitted = true;
      h->workers.push_back(wk);
    }
    std::vector<int32_t> rcs(K, B200BO_OK);
    std::vector<std::thread> th;
    int64_t launches0 = 0;
    for (int k = 0; k < K; ++k) launches0 += h->workers[k]->launches;
    auto run = [&](int k) {
      b200bo_handle_s* wk = h->workers[k];
      cudaSetDevice(wk->device);
      if (wk->synced_version != h->data_version) {
        wk->hX = h->hX; wk->hy = h->hy; wk->N = h->N;
        const int32_t e = upload_data(wk);
        if (e) { rcs[k] = e; return; }
        wk->synced_version = h->data_version;
      }
      wk->hp = h->hp; wk->syrk_engine = h->syrk_engine; wk->chol_sched = h->chol_sched; wk->fitted = false;
      wk->chol_graph = K > 1 ? 0 : h->chol_graph;      // several factorisations in flight: graph launches of different workers serialise (measured 0.108 vs 0.068 s); eager look-ahead overlaps them
      int64_t lo, hi; shard_bounds(S, K, k, &lo, &hi);
      if (hi > lo) rcs[k] = b200bo_mll_sweep(wk, Theta + lo * P, P, (int32_t)(hi - lo), mask, mll + lo, dmll ? dmll + lo * P : nullptr);
    };
    const auto t0 = std::chrono::steady_clock::now();
    for (int k = 1; k < K; ++k) th.emplace_back(run, k);
    run(0);
    for (auto& t : th) t.join();
    h->timing[B200BO_T_MLL] = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    for (int k = 0; k < K; ++k) h->launches += h->workers[k]->launches;
    h->launches -= launches0;
    for (int k = 0; k < K; ++k) if (rcs[k]) return fail(h, rcs[k], "sweep worker " + std::to_string(k) + ": " + h->workers[k]->err);
    return B200BO_OK;
  }
  const Hyper saved = h->hp;
  int32_t rc = B200BO_OK;
  cudaEventRecord(h->ev[6], h->stream);
  for (int s = 0; s < S && rc == B200BO_OK; ++s) {
    const double* th = Theta + (int64_t)s * P;
    int i = 0;
    if (m_noise) h->hp.lognoise = th[i++];
    if (m_mean) h->hp.beta = th[i++];
    if (m_kern) { for (auto& l : h->hp.ll) l = th[i++]; h->hp.lsigma = th[i++]; }
    rc = refit(h);
    if (rc) break;
    mll[s] = h->mll;
    if (dmll) {
      double raw[35];
      static const bool by_solves = getenv("B200BO_KINV_SOLVE") != nullptr;   // developer cross-check of the two Sigma^-1 paths
      const bool solves = by_solves && !h->lite;            // workers carry no solve panels
      cudaError_t e = solves ? launch_kinv_solve(h) : launch_kinv(h);
      if (e == cudaSuccess) e = launch_dmll(h, mask, h->dscal + 8, solves ? h->dV : h->dKi);
      if (e == cudaSuccess) e = cudaMemcpyAsync(raw, h->dscal + 8, sizeof(raw), cudaMemcpyDeviceToHost, h->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
      if (e != cudaSuccess) { rc = fail(h, B200BO_ERR_CUDA, std::string("mll gradient: ") + cudaGetErrorString(e)); break; }
      double* g = dmll + (int64_t)s * P;
      int j = 0;
      if (m_noise) g[j++] = exp(2.0 * h->hp.lognoise) * raw[33];
      if (m_mean) g[j++] = raw[34];
      if (m_kern) {
        if (h->iso) { double t = 0.0; for (int d = 0; d < h->D; ++d) t += raw[d]; g[j++] = 0.5 * t; }
        else for (int d = 0; d < h->D; ++d) g[j++] = 0.5 * raw[d];
        g[j++] = raw[32];
      }
    }
  }
  cudaEventRecord(h->ev[7], h->stream);
  cudaStreamSynchronize(h->stream);
  cudaEventElapsedTime(&h->timing[B200BO_T_MLL], h->ev[6], h->ev[7]);
  h->hp = saved;
  h->fitted = false;   // the factor now belongs to the last swept setting; refactor lazily with the restored params
  return rc;
}

B200BO_API int32_t b200bo_last_timing_ms(b200bo_handle_t h, int32_t which, float* ms) {
  if (!h || !ms || which < 0 || which >= B200BO_T_COUNT) return fail(h, B200BO_ERR_ARG, "bad arguments to last_timing_ms");
  if (which == B200BO_T_ACQ_GEMM) {
    cudaSetDevice(h->device);
    float sum = 0.f;
    if (h->gemm_ev_used > 0 && cudaEventSynchronize(h->gemm_ev[h->gemm_ev_used - 1]) == cudaSuccess)
      for (int i = 0; i + 1 < h->gemm_ev_used; i += 2) { float t = 0.f; cudaEventElapsedTime(&t, h->gemm_ev[i], h->gemm_ev[i + 1]); sum += t; }
    h->timing[which] = sum;
  }
  if (which == B200BO_T_ACQ) {   // the _dev entries do not synchronise; resolve the event pair on demand
    cudaSetDevice(h->device);
    if (cudaEventSynchronize(h->ev[5]) == cudaSuccess) cudaEventElapsedTime(&h->timing[B200BO_T_ACQ], h->ev[4], h->ev[5]);
  }
  *ms = h->timing[which];
  return B200BO_OK;
}

B200BO_API int32_t b200bo_fp64_peak_tflops(b200bo_handle_t h, double* tflops) {
  if (!h || !tflops) return fail(h, B200BO_ERR_ARG, "null argument");
  cudaSetDevice(h->device);
  CU(launch_dmma_peak(h, tflops));
  return B200BO_OK;
}

B200BO_API int32_t b200bo_set_syrk_engine(b200bo_handle_t h, int32_t engine) {
  if (!h || engine < 0 || engine > 2) return fail(h, B200BO_ERR_ARG, "engine must be 0 (DMMA), 1 (tcgen05, 128 x 64 tiles) or 2 (tcgen05, 128 x 128 tiles, two passes)");
  h->syrk_engine = engine;
  for (auto* r : h->replicas) r->syrk_engine = engine;
  return B200BO_OK;
}

B200BO_API int32_t b200bo_set_acq_engine(b200bo_handle_t h, int32_t engine) {
  if (!h || engine < 0 || engine > 1) return fail(h, B200BO_ERR_ARG, "engine must be 0 (DMMA blocked solve) or 1 (tcgen05 int8-slice product)");
  h->acq_engine = engine;
  for (auto* r : h->replicas) r->acq_engine = engine;
  return B200BO_OK;
}

B200BO_API int32_t b200bo_i8_peak_tops(b200bo_handle_t h, double* tops) {
  if (!h || !tops) return fail(h, B200BO_ERR_ARG, "null argument");
  cudaSetDevice(h->device);
  CU(launch_i8_peak(h, tops));
  return B200BO_OK;
}

B200BO_API int32_t b200bo_set_knob(b200bo_handle_t h, const char* name, int64_t value) {
  if (!h || !name) return fail(h, B200BO_ERR_ARG, "null argument");
  const std::string k(name);
  if (k == "acq_lanes") { if (value < 1 || value > 2) return fail(h, B200BO_ERR_ARG, "acq_lanes must be 1 or 2"); h->acq_lanes = (int)value; }
  else if (k == "acq_chunk_mb") { if (value < 0) return fail(h, B200BO_ERR_ARG, "acq_chunk_mb must be >= 0"); h->acq_chunk_mb = value; }
  else if (k == "acq_gemm_timing") h->acq_time_gemm = value != 0;
  else if (k == "sweep_workers") { if (value < 0 || value > 32) return fail(h, B200BO_ERR_ARG, "sweep_workers must be in 0..32"); h->sweep_workers = (int)value; }
  else if (k == "chol_sched") { if (value < -1 || value > 2) return fail(h, B200BO_ERR_ARG, "chol_sched must be -1 (default), 0 (in-order), 1 (look-ahead, fused head) or 2 (look-ahead, tile-GEMM heads)"); h->chol_sched = (int)value; for (auto* w : h->workers) w->chol_sched = (int)value; }
  else if (k == "chol_graph") { if (value < -1 || value > 1000) return fail(h, B200BO_ERR_ARG, "chol_graph must be -1 / 1 (default: capture a shape at its 6th consecutive factorisation), 0 (eager launches) or k >= 2 (capture at the k-th)"); h->chol_graph = (int)value; for (auto* w : h->workers) w->chol_graph = (int)value; }
  else return fail(h, B200BO_ERR_ARG, "unknown knob: " + k);
  for (auto* r : h->replicas) b200bo_set_knob(r, name, value);
  return B200BO_OK;
}

B200BO_API int32_t b200bo_launch_count(b200bo_handle_t h, int64_t* n) {
  if (!h || !n) return fail(h, B200BO_ERR_ARG, "null argument");
  *n = h->launches;
  return B200BO_OK;
}

}  // extern "C"
