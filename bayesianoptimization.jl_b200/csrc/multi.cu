// multi.cu -- multi-GPU behind the C ABI (reference loop: src/acquisition.jl:58-66, a plain running max over independent starts).
//
// The M candidate columns shard over the GPUs in contiguous blocks; X, y and the hyper-parameters go to every GPU and the factor is
// recomputed redundantly (identical kernels on identical inputs: identical bits).  The one exchange step is ONE ncclAllGather of a
// fixed-size record per rank -- (best value, global index, the winning point) -- followed by a deterministic merge (largest value, then
// lowest global index; a rank with no winner carries index -1) in a one-thread kernel, so every rank ends with the same global best.
// Two ways to get the ranks:
//   b200bo_create_multi      ONE process drives n_gpus replicas (what a Julia host uses): ncclCommInitAll, one host thread per replica
//                            for the blocking model updates, the all-gather as one ncclGroup over the replicas' streams
//   b200bo_comm_init_rank    one process per GPU (torchrun): the caller distributes the ncclUniqueId, every rank attaches a
//                            communicator to its own handle and b200bo_acquire(_dev) returns the GLOBAL best
// NCCL is loaded with dlopen at first use (libnccl.so.2: the copy already in the process -- e.g. PyTorch's -- or the system one), so
// the single-GPU library has no link-time dependency on it.
#include <dlfcn.h>
#include <nccl.h>
#include <cmath>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include "common.cuh"
#include "handle.h"

namespace b200bo {

namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
};

NcclApi& nccl() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
    api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) { api.err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return api; }
#define B200BO_SYM(field, sym)                                                              \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, sym));                  \
  if (!api.field) { api.err = std::string("NCCL symbol missing: ") + sym; api.lib = nullptr; return api; }
  B200BO_SYM(GetUniqueId, "ncclGetUniqueId")
  B200BO_SYM(CommInitRank, "ncclCommInitRank")
  B200BO_SYM(CommInitAll, "ncclCommInitAll")
  B200BO_SYM(CommDestroy, "ncclCommDestroy")
  B200BO_SYM(AllGather, "ncclAllGather")
  B200BO_SYM(GroupStart, "ncclGroupStart")
  B200BO_SYM(GroupEnd, "ncclGroupEnd")
  B200BO_SYM(GetErrorString, "ncclGetErrorString")
#undef B200BO_SYM
  return api;
}

// record[0] = best value, record[1] = global index (bit pattern of an int64), record[2 .. 2+D) = the winning point
__global__ void pack_best_kernel(const b200bo_best_t* __restrict__ best, const double* __restrict__ Xs, int64_t idx_offset, int D,
                                 double* __restrict__ rec) {
  const int t = threadIdx.x;
  const long long ix = best->index;
  if (t == 0) { rec[0] = best->value; rec[1] = __longlong_as_double(ix); }
  if (t < D) rec[2 + t] = ix >= 0 ? Xs[(ix - idx_offset) * D + t] : 0.0;
}

// deterministic merge of the gathered records: largest value, then lowest global index; index < 0 and NaN never win
__global__ void merge_best_kernel(const double* __restrict__ gathered, int world, int D, b200bo_best_t* __restrict__ out, double* __restrict__ out_x) {
  __shared__ int win;
  if (threadIdx.x == 0) {
    double bv = -INFINITY; long long bi = -1; int w = -1;
    for (int r = 0; r < world; ++r) {
      const double v = gathered[(size_t)r * B200BO_REC_DOUBLES];
      const long long ix = __double_as_longlong(gathered[(size_t)r * B200BO_REC_DOUBLES + 1]);
      if (ix >= 0 && v == v && (bi < 0 || v > bv || (v == bv && ix < bi))) { bv = v; bi = ix; w = r; }
    }
    out->value = bv; out->index = bi;
    win = w;
  }
  __syncthreads();
  if (out_x && (int)threadIdx.x < D) out_x[threadIdx.x] = win >= 0 ? gathered[(size_t)win * B200BO_REC_DOUBLES + 2 + threadIdx.x] : 0.0;
}

}  // namespace

const char* nccl_load_error() { return nccl().lib ? nullptr : nccl().err.c_str(); }

cudaError_t comm_buffers(b200bo_handle_s* h, int world) {
  if (h->drec && h->rec_world >= world) return cudaSuccess;
  cudaFree(h->drec);
  h->drec = nullptr; h->rec_world = 0;
  // [record of this rank | gathered records of all ranks | merged best (16 B) | merged point (32 doubles)]
  cudaError_t e = cudaMalloc(&h->drec, sizeof(double) * (size_t)B200BO_REC_DOUBLES * (size_t)(world + 2));
  if (e == cudaSuccess) h->rec_world = world;
  return e;
}

// this rank's record from (dbest, dXs); enqueue only
cudaError_t launch_pack_best(b200bo_handle_s* h, const b200bo_best_t* dbest, const double* dXs, int64_t idx_offset) {
  pack_best_kernel<<<1, 32, 0, h->stream>>>(dbest, dXs, idx_offset, h->D, h->drec);
  h->launches++;
  return cudaGetLastError();
}

cudaError_t launch_merge_best(b200bo_handle_s* h, int world, b200bo_best_t* dout, double* dout_x) {
  merge_best_kernel<<<1, 32, 0, h->stream>>>(h->drec + B200BO_REC_DOUBLES, world, h->D, dout, dout_x);
  h->launches++;
  return cudaGetLastError();
}

// the ONE exchange step of a rank that owns a communicator (one process per GPU)
int nccl_allgather_records(b200bo_handle_s* h, std::string* err) {
  NcclApi& n = nccl();
  if (!n.lib) { *err = n.err; return -1; }
  const ncclResult_t r = n.AllGather(h->drec, h->drec + B200BO_REC_DOUBLES, sizeof(double) * B200BO_REC_DOUBLES, ncclChar,
                                     static_cast<ncclComm_t>(h->comm), h->stream);
  if (r != ncclSuccess) { *err = std::string("ncclAllGather: ") + n.GetErrorString(r); return -1; }
  return 0;
}

// single process, all replicas: one group call, each on its replica's stream
int nccl_allgather_group(const std::vector<b200bo_handle_s*>& reps, std::string* err) {
  NcclApi& n = nccl();
  if (!n.lib) { *err = n.err; return -1; }
  ncclResult_t r = n.GroupStart();
  for (size_t i = 0; i < reps.size() && r == ncclSuccess; ++i) {
    b200bo_handle_s* h = reps[i];
    cudaSetDevice(h->device);
    r = n.AllGather(h->drec, h->drec + B200BO_REC_DOUBLES, sizeof(double) * B200BO_REC_DOUBLES, ncclChar, static_cast<ncclComm_t>(h->comm), h->stream);
  }
  const ncclResult_t r2 = n.GroupEnd();
  if (r == ncclSuccess) r = r2;
  if (r != ncclSuccess) { *err = std::string("ncclAllGather (group): ") + n.GetErrorString(r); return -1; }
  return 0;
}

int nccl_unique_id(uint8_t* id128, std::string* err) {
  NcclApi& n = nccl();
  if (!n.lib) { *err = n.err; return -1; }
  ncclUniqueId id;
  const ncclResult_t r = n.GetUniqueId(&id);
  if (r != ncclSuccess) { *err = std::string("ncclGetUniqueId: ") + n.GetErrorString(r); return -1; }
  memcpy(id128, id.internal, NCCL_UNIQUE_ID_BYTES);
  return 0;
}

int nccl_init_rank(b200bo_handle_s* h, int world, int rank, const uint8_t* id128, std::string* err) {
  NcclApi& n = nccl();
  if (!n.lib) { *err = n.err; return -1; }
  ncclUniqueId id;
  memcpy(id.internal, id128, NCCL_UNIQUE_ID_BYTES);
  ncclComm_t c = nullptr;
  cudaSetDevice(h->device);
  const ncclResult_t r = n.CommInitRank(&c, world, id, rank);
  if (r != ncclSuccess) { *err = std::string("ncclCommInitRank: ") + n.GetErrorString(r); return -1; }
  h->comm = c; h->comm_world = world; h->comm_rank = rank;
  return 0;
}

int nccl_init_all(const std::vector<b200bo_handle_s*>& reps, std::string* err) {
  NcclApi& n = nccl();
  if (!n.lib) { *err = n.err; return -1; }
  std::vector<int> devs;
  for (auto* h : reps) devs.push_back(h->device);
  std::vector<ncclComm_t> comms(reps.size(), nullptr);
  const ncclResult_t r = n.CommInitAll(comms.data(), (int)reps.size(), devs.data());
  if (r != ncclSuccess) { *err = std::string("ncclCommInitAll: ") + n.GetErrorString(r); return -1; }
  for (size_t i = 0; i < reps.size(); ++i) { reps[i]->comm = comms[i]; reps[i]->comm_world = (int)reps.size(); reps[i]->comm_rank = (int)i; }
  return 0;
}

void nccl_destroy(b200bo_handle_s* h) {
  if (h->comm && nccl().lib) nccl().CommDestroy(static_cast<ncclComm_t>(h->comm));
  h->comm = nullptr; h->comm_world = 1; h->comm_rank = 0;
}

// contiguous block [lo, hi) of `total` units owned by part r of R (remainder spread over the low parts)
void shard_bounds(int64_t total, int R, int r, int64_t* lo, int64_t* hi) {
  const int64_t base = total / R, rem = total % R;
  *lo = r * base + (r < rem ? r : rem);
  *hi = *lo + base + (r < rem ? 1 : 0);
}

}  // namespace b200bo
