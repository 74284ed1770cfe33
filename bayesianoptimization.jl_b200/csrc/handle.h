// handle.h -- host-side state of one GP model resident on one B200, and the launcher prototypes of the kernels.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/b200bo.h"

namespace b200bo {

struct Hyper {               // current hyper-parameters (EXT GaussianProcesses.jl parametrisation)
  double lognoise = -2.0;    // sigma_n = exp(lognoise)
  double beta = 0.0;         // MeanConst
  double lsigma = 0.0;       // sigma_f = exp(lsigma)
  std::vector<double> ll;    // log length-scales: 1 (Iso) or D (Ard)
};

}  // namespace b200bo

#define B200BO_REC_DOUBLES 34   // exchange record of one rank: best value, global index (int64 bits), the winning point (D <= 32)

struct b200bo_handle_s {
  int device = 0, D = 0, kernel_kind = 0, mean_kind = 0, fam = 0;
  bool iso = false;
  int num_sms = 148;
  int64_t cap = 0;           // capacity in points, multiple of NB
  int64_t N = 0, Np = 0;     // observations, padded to a multiple of NB
  int64_t ld = 0;            // leading dimension of the factor (== cap)
  b200bo::Hyper hp;
  std::vector<double> hX, hy;     // host copies of model.x (D x N col-major), model.y
  // device buffers
  double* dX = nullptr;      // [cap][D] raw inputs (point-major == D x N column-major)
  double* dZ = nullptr;      // [cap][D] inputs scaled by 1/l_d
  double* dZk = nullptr;     // [cap/64] K1 staging blocks (kmat.cu): DMMA fragment order of z for 64 points as tile rows / cols + |z|^2/2
  double* dy = nullptr;      // [cap]
  double* dw = nullptr;      // [cap] work vector for the single-RHS solves
  double* dalpha = nullptr;  // [cap] (zero in the padding)
  double* dz = nullptr;      // [cap] z = L^-1 (y - m), kept for elastic appends
  int* dflags = nullptr;     // [cap/128] "alpha block ready" epochs of the one-launch backward solve (solve.cu)
  int solve_epoch = 0;
  double noise_total = 0.0;  // diagonal noise of the current factor (incl. make_posdef! jitter)
  double* dinv_ell = nullptr;  // [D]
  double* dL = nullptr;      // [ld][ld] mirrored factor: lower = L (row-major) == U column-major, upper = L^T
  double* dLinv = nullptr;   // [cap/NB][NB][NB] inverses of the diagonal blocks
  double* dLinvT = nullptr;  //   and their transposes
  double* dV = nullptr;      // per-CTA solve panels [nslots][TILE_N][ld]
  double *dKi = nullptr, *dWT = nullptr, *dTT = nullptr;   // [cap][cap] each, allocated by the first MAP gradient (kinv.cu)
  CUtensorMap tmKi64, tmWT128, tmWT64, tmTT128;
  void* dSl = nullptr;       // [8 slices][cap][512] int8 slices of the current outer panel (syrk_i8.cu), allocated on first use
  double* dSe = nullptr;     // [cap] per-row scale 2^(e-6) of the slices
  int i8_nkb = 4;            // 128-column k-blocks in the current slices (= inner panels per outer panel)
  CUtensorMap tmSlA, tmSlB;
  // K6 on tcgen05 (acq_i8.cu): int8 slices of W = L^-1 and (gradient) Sigma^-1 of the current factor, [7][cap][cap] + row scales; the
  // per-chunk k* slices [7][CH][Np], partial sums, w = Sigma^-1 k* and per-block bests
  void *dWs = nullptr, *dKs = nullptr, *dBs = nullptr;
  double *dWe = nullptr, *dKe = nullptr, *dMuP = nullptr, *dWg = nullptr;
  b200bo_best_t* dcta_best2 = nullptr;
  size_t bs_bytes = 0, part_bytes = 0, wg_bytes = 0;
  int64_t bs_np = 0, bs_ch = 0, nbest2 = 0;
  CUtensorMap tmWsB, tmKsB, tmBsA[4];   // tmBsA: k* slices of the two lanes, then the v slices of the two lanes
  bool bs_grad = false, wt_valid = false;   // v-slice buffers / maps exist; h->dWT holds W^T of the CURRENT factor
  cudaEvent_t acq_ev[2] = {nullptr, nullptr};   // fork / join of the two chunk lanes
  int acq_lanes = 2;         // chunk lanes of the tcgen05 acquisition path (1: everything on the handle's stream)
  int64_t acq_chunk_mb = 0;  // 0: default L2 budget per chunk
  bool acq_time_gemm = false;   // time every slice-product launch with CUDA events (forces one lane): B200BO_T_ACQ_GEMM
  std::vector<cudaEvent_t> gemm_ev;
  int gemm_ev_used = 0;
  int acq_engine = -1;       // -1: default (tcgen05 unless B200BO_ACQ_I8=0), 0: DMMA solve (acq.cu), 1: tcgen05 int8-slice GEMM (acq_i8.cu)
  int64_t nslots = 0;
  double* dscal = nullptr;   // small scalar outputs (logdet, r'alpha, ...)
  int* dinfo = nullptr;      // non-PD flag
  b200bo_best_t* dcta_best = nullptr;   // [grid]
  b200bo_best_t* dbest = nullptr;
  double* dpart = nullptr;   // partial sums for the mll gradient
  // TMA descriptors (128B-swizzled 16-double-wide boxes) over the factor, the solve panels and the inverted diagonal blocks
  CUtensorMap tmL, tmL64, tmV, tmLinv, tmLinvT;
  double* dio = nullptr;     // staging for host-pointer entry points
  double* dlbub = nullptr;   // [2][D] box bounds of the search
  int64_t dio_bytes = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  cudaStream_t stream2 = nullptr;     // look-ahead stream of the factorisation (trailing update k overlaps panel k+1)
  cudaStream_t stream3 = nullptr;     // rest stream of the look-ahead schedule (chol.cu): everything of panel k that potrf(k+1) does not need
  cudaStream_t stream4 = nullptr, stream5 = nullptr;   //   its bulk-update and rider streams
  std::vector<cudaEvent_t> ch_ev;     // per block: potrf / head / column path / panel solve / bulk update / boundary share done
  double* dD = nullptr;               // [128][128] scratch accumulator of the boundary diagonal block (look-ahead schedule)
  cudaGraphExec_t chol_graph_exec = nullptr;   // the captured look-ahead factorisation of the current shape (chol.cu: launch_cholesky)
  uint64_t chol_graph_key = 0, chol_seen_key = 0;
  int chol_seen_count = 0;            // consecutive factorisations of the shape chol_seen_key
  int64_t chol_graph_launches = 0;
  int chol_graph_syrk_ev = 0;
  int chol_graph = -1;                // -1 / 1: default (capture a shape at its 6th consecutive factorisation), 0: eager launches, k >= 2: capture at the k-th
  int chol_sched = -1;                // -1: default (look-ahead), 0: in-order schedule of round 1, 1: look-ahead
  std::vector<cudaEvent_t> la_ev;     // look-ahead dependencies
  std::vector<cudaEvent_t> fw_ev;     // per panel: the forward solve of y - m rides along the factorisation on stream2
  cudaEvent_t ev[8] = {};
  std::vector<cudaEvent_t> syrk_ev;   // start/stop pairs around every trailing-update launch of the last factorisation
  int syrk_ev_used = 0;
  int syrk_engine = -1;      // -1: default (tcgen05 unless B200BO_SYRK_I8=0), 0: DMMA, 1: tcgen05
  // multi-GPU (multi.cu): a communicator over the ranks of this model (one process per GPU: b200bo_comm_init_rank; one process for all
  // GPUs: b200bo_create_multi, whose parent handle owns one child replica per further device)
  void* comm = nullptr;      // ncclComm_t
  int comm_world = 1, comm_rank = 0;
  double* drec = nullptr;    // [own record | gathered records | merged best + point]
  int rec_world = 0;
  std::vector<b200bo_handle_s*> replicas;   // children of a multi handle (the parent itself is rank 0)
  bool is_replica = false;   // a child of a multi handle
  // MAP sweeps (b200bo_mll_sweep) run on worker models of the same device, several settings in flight at once, each on its own streams
  // and buffers: one setting's chain of diagonal blocks hides behind another's tile GEMMs, and the parent's own factor stays valid
  std::vector<b200bo_handle_s*> workers;
  b200bo_handle_s* joint = nullptr;      // worker model of the joint posterior sample (b200bo_rand_joint): observations + sample points
  int64_t data_version = 0, synced_version = -1;
  bool lite = false;         // a worker: no acquisition solve panels
  int sweep_workers = 6;
  bool in_multi = false;     // inside a fan-out of the parent: behave as a single-GPU handle (no exchange)
  std::vector<int32_t> prior_kind;      // per parameter (full order): 0 flat, 1 Normal(prior_a, prior_b); empty = all flat
  std::vector<double> prior_a, prior_b;
  bool fitted = false;
  bool need_upload = false;  // device copies of X / y are stale (a failed elastic append): re-upload before the next refactor
  int acq_ready = 0;         // bit 0: W = L^-1 sliced for the tcgen05 acquisition path, bit 1: Sigma^-1 sliced (acq_i8.cu)
  double mll = 0.0;
  int jitter = 0;
  int64_t launches = 0;
  float timing[B200BO_T_COUNT] = {};
  std::string err;
};

namespace b200bo {

// tmap.cu
cudaError_t make_tensor_maps(b200bo_handle_s* h);
// kmat.cu
cudaError_t launch_scale_inputs(b200bo_handle_s* h, int64_t n0, int64_t n1);
cudaError_t launch_kmat(b200bo_handle_s* h, double* dK, int64_t ld, int64_t N, int64_t Np, double noise, bool pad_identity);
// chol.cu
cudaError_t launch_cholesky(b200bo_handle_s* h);   // in place on h->dL (lower triangle), fills dLinv/dLinvT, upper mirror
void release_cholesky_graph(b200bo_handle_s* h);
// syrk_i8.cu
bool syrk_i8_enabled();
cudaError_t launch_slice_panel(b200bo_handle_s* h, cudaStream_t st, int row0, int col0, int nkb);
cudaError_t launch_syrk_i8(b200bo_handle_s* h, cudaStream_t st, int bi_lo, int col2_lo, int col2_hi, int* ntiles, int max_ctas);
// solve.cu
cudaError_t launch_alpha_mll(b200bo_handle_s* h, bool have_z);  // dw = y - m -> dz (unless have_z), dalpha, dscal[0] = logdet, dscal[1] = r'alpha
cudaError_t launch_residual(b200bo_handle_s* h, cudaStream_t st);            // dw = y - m (zero in the padding)
cudaError_t launch_fwd_step(b200bo_handle_s* h, cudaStream_t st, int i, int nblk);   // step i of z = L^-1 dw (i = -1 .. nblk-2)
cudaError_t launch_forward_solve(b200bo_handle_s* h, double* w, double* z, int nblk);
cudaError_t launch_backward_solve(b200bo_handle_s* h, const double* z, double* w, double* alpha, int nblk);
cudaError_t launch_logdet_dot(b200bo_handle_s* h);
// append.cu
cudaError_t launch_append_one(b200bo_handle_s* h, double noise, bool last);
cudaError_t launch_append_block(b200bo_handle_s* h, double noise, int m);   // m <= 16 points inside one 128-block, from W = L^-1 (h->wt_valid)
// acq.cu
struct AcqLaunch {
  int acq_kind = -1;         // -1: predict only
  double p0 = 0.0, p1 = 0.0;
  uint64_t seed = 0;
  int64_t idx_offset = 0;
  const double* dXs = nullptr;
  int64_t M = 0;
  double *dvalues = nullptr, *dgrad = nullptr, *dmu = nullptr, *dvar = nullptr;
  b200bo_best_t* dbest = nullptr;
  // host mirrors (host-pointer entries): when hXs is set the candidates are NOT on the device yet -- the tcgen05 engine copies chunk c in
  // on its stream lane right before the chunk's kernels and its outputs back right behind them, so the PCIe transfers of one chunk run
  // under the other lane's kernels (the DMMA engine copies everything up front / at the end)
  const double* hXs = nullptr;
  double *hvalues = nullptr, *hgrad = nullptr, *hmu = nullptr, *hvar = nullptr;
};
cudaError_t launch_acquire(b200bo_handle_s* h, const AcqLaunch& a);       // dispatches on h->acq_engine
cudaError_t launch_acquire_i8(b200bo_handle_s* h, const AcqLaunch& a);    // acq_i8.cu
bool acq_i8_default();
size_t acq_smem_bytes(int D);
// search.cu
cudaError_t launch_lhs(b200bo_handle_s* h, double* dXs, int64_t n_total, int64_t offset, int64_t n_local, unsigned long long seed,
                       const double* d_lbub);
cudaError_t launch_ascent(b200bo_handle_s* h, const AcqLaunch& base, double* dX, double* dwork, const double* d_lbub, int steps, double s0);
// multi.cu
const char* nccl_load_error();
cudaError_t comm_buffers(b200bo_handle_s* h, int world);
cudaError_t launch_pack_best(b200bo_handle_s* h, const b200bo_best_t* dbest, const double* dXs, int64_t idx_offset);
cudaError_t launch_merge_best(b200bo_handle_s* h, int world, b200bo_best_t* dout, double* dout_x);
int nccl_allgather_records(b200bo_handle_s* h, std::string* err);
int nccl_allgather_group(const std::vector<b200bo_handle_s*>& reps, std::string* err);
int nccl_unique_id(uint8_t* id128, std::string* err);
int nccl_init_rank(b200bo_handle_s* h, int world, int rank, const uint8_t* id128, std::string* err);
int nccl_init_all(const std::vector<b200bo_handle_s*>& reps, std::string* err);
void nccl_destroy(b200bo_handle_s* h);
void shard_bounds(int64_t total, int R, int r, int64_t* lo, int64_t* hi);
struct LbfgsOpts;
cudaError_t launch_sobol(b200bo_handle_s* h, double* dXs, unsigned long long index0, int64_t n, const double* d_lbub);
cudaError_t launch_lbfgs(b200bo_handle_s* h, const AcqLaunch& base, double* dXe, double* dwork, const double* d_lbub, const LbfgsOpts& o,
                         double maxtime_s, int* rounds_out);
// joint.cu
cudaError_t launch_joint_diag(b200bo_handle_s* h, int64_t n0, int64_t m, double value);
cudaError_t launch_joint_sample(b200bo_handle_s* h, int64_t n0, int64_t m, const double* dmu, double* deps, double* dout,
                                unsigned long long seed, int64_t idx_offset);
// peak.cu
cudaError_t launch_dmma_peak(b200bo_handle_s* h, double* tflops);
cudaError_t launch_i8_peak(b200bo_handle_s* h, double* tops);
// mll.cu
cudaError_t launch_kinv(b200bo_handle_s* h);       // kinv.cu: Sigma^-1 into h->dKi by recursive block inversion + W^T W
cudaError_t launch_linv(b200bo_handle_s* h);       //   first half: W = L^-1 into h->dKi (lower blocks), W^T into h->dWT
cudaError_t launch_kinv_syrk(b200bo_handle_s* h);  //   second half: Sigma^-1 = W^T W into h->dKi
cudaError_t launch_kinv_solve(b200bo_handle_s* h); // acq.cu (MODE 1): the same by column solves against I, into h->dV; cross-check only
cudaError_t launch_dmll(b200bo_handle_s* h, int mask, double* dout /*P*/, const double* Kinv);

}  // namespace b200bo
