// search.cu -- the "next" rows of the search: candidates are born and refined in HBM.
//
//  * lhs_kernel: latin_hypercube_sampling (reference src/utils.jl:101-120) on device.  Per dimension the n strata are
//    assigned by a keyed, stateless pseudo-random PERMUTATION of [0, n) (invertible mix on ceil(log2 n) bits + cycle walking)
//    instead of Random.shuffle!, and jittered by a Philox draw keyed by (seed, dim, global column): every stratum is used exactly
//    once per dimension (the LHS property), columns can be generated in any order / on any rank (sharding-invariant).
//  * ascent_step_kernel: M projected-gradient ascents in lock-step on the fused value+gradient kernel -- what one NLopt run per
//    restart does in the reference (src/acquisition.jl:59; box bounds :28-29), with a per-candidate adaptive step.
#include <chrono>
#include "common.cuh"
#include "lbfgs.cuh"
#include "sobol_dirs.cuh"
#include "handle.h"

namespace b200bo {

// ---- Philox4x32-10 (same rounds as acq.cu) returning the 4 words ------------------------------------------------------
__device__ __host__ inline void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// invertible mix on `bits` bits (odd multiplies, xor-shifts, adds), iterated until the image falls in [0, n)
__device__ __host__ inline uint64_t perm_index(uint64_t x, uint64_t n, int bits, const uint32_t key[4]) {
  const uint64_t mask = (bits >= 64) ? ~0ull : ((1ull << bits) - 1ull);
  const int sh = bits / 2 + 1;
  do {
    x = (x + key[0]) & mask;
    x = (x * (2ull * key[1] + 1ull)) & mask;
    x ^= x >> sh;
    x = (x * (2ull * key[2] + 1ull)) & mask;
    x ^= x >> sh;
    x = (x + key[3]) & mask;
  } while (x >= n);
  return x;
}

__global__ void lhs_kernel(double* __restrict__ Xs, int D, int64_t n_total, int64_t offset, int64_t n_local, int bits,
                           const double* __restrict__ lbub, unsigned long long seed) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_local * D) return;
  const int64_t jl = e / D;
  const int d = (int)(e - jl * D);
  const uint64_t j = (uint64_t)(offset + jl);
  uint32_t key[4], u[4];
  philox4x32((uint32_t)d, 0x4C485321u, 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), key);          // per-dimension permutation key
  const uint64_t stratum = perm_index(j, (uint64_t)n_total, bits, key);
  philox4x32((uint32_t)j, (uint32_t)(j >> 32), (uint32_t)d, 0x4A495454u, (uint32_t)seed, (uint32_t)(seed >> 32), u);   // jitter
  const double jit = (double)((((uint64_t)u[0] << 32) | u[1]) >> 11) * 1.1102230246251565e-16;      // [0, 1)
  const double lo = lbub[d], hi = lbub[D + d];
  const double step = (hi - lo) / (double)n_total;
  Xs[e] = __dadd_rn(lo, __dmul_rn(step, (double)stratum + jit));     // no FMA contraction: bit-equal to the host restatement
}

cudaError_t launch_lhs(b200bo_handle_s* h, double* dXs, int64_t n_total, int64_t offset, int64_t n_local, unsigned long long seed,
                       const double* d_lbub) {
  if (n_local <= 0) return cudaSuccess;
  int bits = 1;
  while (bits < 63 && (1ull << bits) < (uint64_t)n_total) ++bits;
  const int64_t total = n_local * h->D;
  lhs_kernel<<<(int)((total + 255) / 256), 256, 0, h->stream>>>(dXs, h->D, n_total, offset, n_local, bits, d_lbub, seed);
  h->launches++;
  return cudaGetLastError();
}

// ---- batched projected-gradient ascent -------------------------------------------------------------------------------
// state per candidate: best point/value/gradient so far and the current step (in units of the box diagonal fraction).
// iter 0 just records the start; afterwards: improve -> accept and grow the step, otherwise shrink and retry from the best.
__global__ void ascent_step_kernel(double* __restrict__ X, const double* __restrict__ val, const double* __restrict__ grad,
                                   double* __restrict__ Xb, double* __restrict__ Fb, double* __restrict__ Gb, double* __restrict__ S,
                                   const double* __restrict__ lbub, int D, int64_t M, int iter, int last, double s0) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const double v = val[i];
  double s = iter == 0 ? s0 : S[i];
  const bool better = iter == 0 || v > Fb[i];                 // NaN never improves
  if (better) {
    Fb[i] = v;
    for (int d = 0; d < D; ++d) { Xb[i * D + d] = X[i * D + d]; Gb[i * D + d] = grad[i * D + d]; }
    if (iter > 0) s = fmin(s * 1.3, 0.5);
  } else {
    s *= 0.35;
  }
  S[i] = s;
  if (last) {
    for (int d = 0; d < D; ++d) X[i * D + d] = Xb[i * D + d];
    return;
  }
  double nrm = 0.0;
  for (int d = 0; d < D; ++d) { const double gu = Gb[i * D + d] * (lbub[D + d] - lbub[d]); nrm = fma(gu, gu, nrm); }
  nrm = sqrt(nrm);
  for (int d = 0; d < D; ++d) {
    const double lo = lbub[d], hi = lbub[D + d], rng = hi - lo;
    double x = Xb[i * D + d];
    if (nrm > 0.0 && nrm == nrm) x += rng * s * (Gb[i * D + d] * rng / nrm);
    X[i * D + d] = fmin(fmax(x, lo), hi);
  }
}

__global__ void argmax_values_kernel(const double* __restrict__ v, int64_t M, int64_t idx_offset, b200bo_best_t* __restrict__ out) {
  __shared__ double sv[256];
  __shared__ long long si[256];
  double bv = -INFINITY; long long bi = -1;
  for (int64_t i = threadIdx.x; i < M; i += 256) {
    const double x = v[i];
    if (x > bv) { bv = x; bi = idx_offset + i; }             // ascending i per thread: first strict maximum
  }
  sv[threadIdx.x] = bv; si[threadIdx.x] = bi;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const double ov = sv[threadIdx.x + o]; const long long oi = si[threadIdx.x + o];
      if (oi >= 0 && (ov > sv[threadIdx.x] || (ov == sv[threadIdx.x] && (si[threadIdx.x] < 0 || oi < si[threadIdx.x])))) { sv[threadIdx.x] = ov; si[threadIdx.x] = oi; }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out->value = sv[0]; out->index = si[0]; }
}

cudaError_t launch_ascent(b200bo_handle_s* h, const AcqLaunch& base, double* dX, double* dwork, const double* d_lbub, int steps, double s0) {
  const int64_t M = base.M, D = h->D;
  double* val = dwork;                 // [M]
  double* grad = val + M;              // [M*D]
  double* Xb = grad + M * D;           // [M*D]
  double* Gb = Xb + M * D;             // [M*D]
  double* Fb = Gb + M * D;             // [M]
  double* S = Fb + M;                  // [M]
  for (int it = 0; it <= steps; ++it) {
    AcqLaunch l = base;
    l.dXs = dX; l.dvalues = val; l.dgrad = grad; l.dmu = nullptr; l.dvar = nullptr; l.dbest = nullptr;
    cudaError_t e = launch_acquire(h, l);
    if (e != cudaSuccess) return e;
    ascent_step_kernel<<<(int)((M + 127) / 128), 128, 0, h->stream>>>(dX, val, grad, Xb, Fb, Gb, S, d_lbub, (int)D, M, it, it == steps ? 1 : 0, s0);
    h->launches++;
  }
  if (base.dvalues) cudaMemcpyAsync(base.dvalues, Fb, sizeof(double) * M, cudaMemcpyDeviceToDevice, h->stream);
  if (base.dbest) {
    argmax_values_kernel<<<1, 256, 0, h->stream>>>(Fb, M, base.idx_offset, base.dbest);
    h->launches++;
  }
  return cudaGetLastError();
}

// ---- Sobol points on device (reference ScaledSobolIterator, src/utils.jl:64-87; EXT Sobol.jl, Joe-Kuo direction numbers) ---------------
// point `index` of the unscrambled sequence in Gray-code order (index 0 = the origin, which Sobol.jl never emits: its k-th point is
// index k), scaled to the box as next!(seq, lb, ub) does.  Stateless: any block of indices on any rank.
__global__ void sobol_kernel(double* __restrict__ Xs, int D, unsigned long long index0, int64_t n, const double* __restrict__ lbub) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * D) return;
  const int64_t j = e / D;
  const int d = (int)(e - j * D);
  const unsigned long long idx = index0 + (unsigned long long)j;
  uint32_t gray = (uint32_t)(idx ^ (idx >> 1)), x = 0u;
  for (int b = 0; gray; ++b, gray >>= 1)
    if (gray & 1u) x ^= SOBOL_V[d][b];
  const double u = (double)x * 2.3283064365386963e-10;         // / 2^32
  const double lo = lbub[d], hi = lbub[D + d];
  Xs[e] = __dadd_rn(lo, __dmul_rn(hi - lo, u));
}

cudaError_t launch_sobol(b200bo_handle_s* h, double* dXs, unsigned long long index0, int64_t n, const double* d_lbub) {
  if (n <= 0) return cudaSuccess;
  const int64_t total = n * h->D;
  sobol_kernel<<<(int)((total + 255) / 256), 256, 0, h->stream>>>(dXs, h->D, index0, n, d_lbub);
  h->launches++;
  return cudaGetLastError();
}

// ---- M box-bounded L-BFGS ascents in lock-step (lbfgs.cuh): one thread per restart consumes the fused launch's value + gradient ----
__global__ void lbfgs_step_kernel(double* __restrict__ state, double* __restrict__ Xe, const double* __restrict__ val, const double* __restrict__ grad,
                                  const double* __restrict__ lbub, int D, int64_t M, LbfgsOpts o, int* __restrict__ running) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  double* st = state + i * lbfgs_state_doubles(D);
  lbfgs_step(st, Xe + i * D, val[i], grad + i * D, lbub, lbub + D, D, o);
  if (LbfgsState(st, D).status() == 0.0) atomicAdd(running, 1);
}

__global__ void lbfgs_finish_kernel(double* __restrict__ state, double* __restrict__ Xe, double* __restrict__ val, double* __restrict__ evals, int D,
                                    int64_t M) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  LbfgsState s(state + i * lbfgs_state_doubles(D), D);
  for (int k = 0; k < D; ++k) Xe[i * D + k] = s.x[k];
  val[i] = s.f();
  if (evals) evals[i] = s.evals();
}

// dXe [M][D]: starts in, maximisers out; dwork: state [M][B] | val [M] | grad [M][D] | evals [M]; returns the number of rounds
cudaError_t launch_lbfgs(b200bo_handle_s* h, const AcqLaunch& base, double* dXe, double* dwork, const double* d_lbub, const LbfgsOpts& o,
                         double maxtime_s, int* rounds_out) {
  const int64_t M = base.M, D = h->D, B = lbfgs_state_doubles((int)D);
  double* state = dwork;
  double* val = state + M * B;
  double* grad = val + M;
  double* evals = grad + M * D;
  cudaError_t e = cudaMemsetAsync(state, 0, sizeof(double) * M * B, h->stream);
  if (e != cudaSuccess) return e;
  int* drun = h->dinfo + 3;
  const auto t0 = std::chrono::steady_clock::now();
  const int cap = o.maxeval > 0 ? o.maxeval : 100000;
  auto enqueue_round = [&]() -> cudaError_t {
    AcqLaunch l = base;
    l.dXs = dXe; l.dvalues = val; l.dgrad = grad; l.dmu = nullptr; l.dvar = nullptr; l.dbest = nullptr;
    cudaError_t er = launch_acquire(h, l);
    if (er != cudaSuccess) return er;
    if ((er = cudaMemsetAsync(drun, 0, sizeof(int), h->stream)) != cudaSuccess) return er;
    lbfgs_step_kernel<<<(int)((M + 127) / 128), 128, 0, h->stream>>>(state, dXe, val, grad, d_lbub, (int)D, M, o, drun);
    h->launches++;
    return cudaGetLastError();
  };
  auto still_running = [&](int* running) -> cudaError_t {
    cudaError_t er = cudaMemcpyAsync(running, drun, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    return er != cudaSuccess ? er : cudaStreamSynchronize(h->stream);
  };
  auto out_of_time = [&]() { return maxtime_s > 0.0 && std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() >= maxtime_s; };   // NLopt maxtime
  // A round is ~7 small launches and the host's look at the running count costs a stream synchronisation: at the sizes of a BO loop
  // (16 starts, N in the hundreds) that latency IS the iteration.  After the first round (eager: it makes every lazy allocation) a block
  // of RB rounds is captured into a CUDA graph and replayed; the host looks at the running count once per block.  Runs that have stopped
  // are no-ops in lbfgs_step (their status is set, per-run maxeval is enforced there), so the extra rounds of a block change nothing.
  constexpr int RB = 8;
  static const bool use_graph = !(getenv("B200BO_LBFGS_GRAPH") && atoi(getenv("B200BO_LBFGS_GRAPH")) == 0);
  int rounds = 0, running = 1;
  if ((e = enqueue_round()) != cudaSuccess) return e;
  if ((e = still_running(&running)) != cudaSuccess) return e;
  ++rounds;
  cudaGraphExec_t exec = nullptr;
  int64_t block_launches = 0;
  cudaStreamCaptureStatus cst = cudaStreamCaptureStatusNone;
  if (running != 0 && rounds < cap && !out_of_time() && use_graph && cudaStreamIsCapturing(h->stream, &cst) == cudaSuccess &&
      cst == cudaStreamCaptureStatusNone && cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
    const int64_t l0 = h->launches;
    for (int r = 0; r < RB && e == cudaSuccess; ++r) e = enqueue_round();
    cudaGraph_t graph = nullptr;
    const cudaError_t e2 = cudaStreamEndCapture(h->stream, &graph);
    block_launches = h->launches - l0;
    h->launches = l0;
    if (e == cudaSuccess && e2 == cudaSuccess && graph && cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) exec = nullptr;
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    e = cudaSuccess;
  }
  while (running != 0 && rounds < cap && !out_of_time()) {
    if (exec) {
      if ((e = cudaGraphLaunch(exec, h->stream)) != cudaSuccess) break;
      h->launches += block_launches;
      rounds += RB;
    } else {
      if ((e = enqueue_round()) != cudaSuccess) break;
      ++rounds;
    }
    if ((e = still_running(&running)) != cudaSuccess) break;
  }
  if (exec) cudaGraphExecDestroy(exec);
  if (e != cudaSuccess) return e;
  lbfgs_finish_kernel<<<(int)((M + 127) / 128), 128, 0, h->stream>>>(state, dXe, val, evals, (int)D, M);
  h->launches++;
  if (base.dvalues) cudaMemcpyAsync(base.dvalues, val, sizeof(double) * M, cudaMemcpyDeviceToDevice, h->stream);
  if (base.dbest) {
    argmax_values_kernel<<<1, 256, 0, h->stream>>>(val, M, base.idx_offset, base.dbest);
    h->launches++;
  }
  if (rounds_out) *rounds_out = rounds;
  return cudaGetLastError();
}

}  // namespace b200bo
