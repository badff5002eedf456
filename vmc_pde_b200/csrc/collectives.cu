// Cross-GPU sums of the TDVP moments for hosts that do not bring torch.distributed (a C or JAX driver): the two reductions a
// right-hand side needs -- packed first moments, packed second moments -- as NCCL all-reduces on the caller's communicator and
// stream.  Replaces the ten host-staged MPI.Allreduce calls of mpi_wrapper.py:129-274 as issued from tdvp.py:37-47.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2: the copy the host process already loaded, e.g. torch's or JAX's, is
// reused), so libvmcpde.so has no link-time NCCL dependency and single-GPU users need no NCCL at all.
//
// The Gram matrices are symmetric and only their upper-triangular 128 x 128 tiles are computed (gram.cu), so what crosses
// NVLink is the PACKED upper tiles: tiles (tiles + 1) / 2 * 128 * 128 doubles per matrix instead of Pp^2 (272 MB instead of
// 537 MB at P = 8187).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstdint>
#include <cstring>
#include <string>
#include "common.cuh"

namespace vmc {

// tile (ti <= tj) -> slot ti * tiles - ti (ti - 1) / 2 + (tj - ti); inside a tile row-major 128 x 128
__global__ void __launch_bounds__(256) pack_tiles_kernel(const double* __restrict__ S, double* __restrict__ packed, int tiles, int Pp, int unpack,
                                                         double* __restrict__ Sout) {
  const int ti = blockIdx.y, tj = blockIdx.x;
  if (tj < ti) return;
  const long long slot = (long long)ti * tiles - (long long)ti * (ti - 1) / 2 + (tj - ti);
  double2* p = reinterpret_cast<double2*>(packed + slot * 16384);
  for (int e = threadIdx.x; e < 8192; e += 256) {
    const int r = e >> 6, c2 = e & 63;
    const long long g = ((long long)(ti * 128 + r) * Pp + tj * 128) / 2 + c2;
    if (unpack) reinterpret_cast<double2*>(Sout)[g] = p[e];
    else p[e] = reinterpret_cast<const double2*>(S)[g];
  }
}

typedef int (*NcclAllReduceFn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*NcclGetUniqueIdFn)(void*);
typedef int (*NcclCommDestroyFn)(void*);
typedef const char* (*NcclGetErrorStringFn)(int);
struct NcclId { char internal[128]; };   // ncclUniqueId (passed by value to ncclCommInitRank)

struct NcclApi {
  void* handle = nullptr;
  NcclAllReduceFn all_reduce = nullptr;
  NcclGetUniqueIdFn get_unique_id = nullptr;
  int (*comm_init_rank)(void**, int, NcclId, int) = nullptr;
  NcclCommDestroyFn comm_destroy = nullptr;
  NcclGetErrorStringFn error_string = nullptr;
};

static int nccl_api(NcclApi** out) {
  // resolved once (thread-safe static initialisation); a failed resolution is remembered and reported on every call
  struct Resolved { NcclApi api; int rc = 0; std::string err; };
  static const Resolved r = [] {
    Resolved q;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      q.api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (q.api.handle) break;
    }
    if (!q.api.handle) {
      const char* why = dlerror();
      q.rc = VMCPDE_EUNSUPPORTED;
      q.err = std::string("NCCL is not loadable (dlopen libnccl.so.2): ") + (why ? why : "unknown error");
      return q;
    }
    q.api.all_reduce = (NcclAllReduceFn)dlsym(q.api.handle, "ncclAllReduce");
    q.api.get_unique_id = (NcclGetUniqueIdFn)dlsym(q.api.handle, "ncclGetUniqueId");
    q.api.comm_init_rank = (int (*)(void**, int, NcclId, int))dlsym(q.api.handle, "ncclCommInitRank");
    q.api.comm_destroy = (NcclCommDestroyFn)dlsym(q.api.handle, "ncclCommDestroy");
    q.api.error_string = (NcclGetErrorStringFn)dlsym(q.api.handle, "ncclGetErrorString");
    if (!q.api.all_reduce || !q.api.get_unique_id || !q.api.comm_init_rank || !q.api.comm_destroy) {
      q.rc = VMCPDE_EUNSUPPORTED;
      q.err = "libnccl.so.2 lacks ncclAllReduce / ncclGetUniqueId / ncclCommInitRank / ncclCommDestroy";
    }
    return q;
  }();
  if (r.rc) return set_error(r.rc, r.err);
  *out = const_cast<NcclApi*>(&r.api);
  return 0;
}

static int nccl_fail(NcclApi* api, const char* what, int rc) {
  return set_error(VMCPDE_ECUDA, std::string(what) + ": " + (api->error_string ? api->error_string(rc) : "NCCL error") + " (" + std::to_string(rc) + ")");
}

}  // namespace vmc

using namespace vmc;

// doubles of the packed upper tiles of one Pp x Pp matrix
extern "C" __attribute__((visibility("default"))) int64_t vmcpde_packed_tiles_len(int32_t Pp) {
  const int64_t t = Pp / 128;
  return t * (t + 1) / 2 * 16384;
}

extern "C" __attribute__((visibility("default"))) int vmcpde_pack_upper_tiles(const double* S, int32_t Pp, double* packed, vmcpde_stream stream) {
  VMC_REQUIRE(S && packed && Pp > 0 && Pp % 128 == 0, "vmcpde_pack_upper_tiles: bad arguments");
  const int t = Pp / 128;
  pack_tiles_kernel<<<dim3(t, t), 256, 0, (cudaStream_t)stream>>>(S, packed, t, Pp, 0, nullptr);
  VMC_LAUNCH_CHECK("pack_tiles_kernel");
  return 0;
}

extern "C" __attribute__((visibility("default"))) int vmcpde_unpack_upper_tiles(const double* packed, int32_t Pp, double* S, vmcpde_stream stream) {
  VMC_REQUIRE(S && packed && Pp > 0 && Pp % 128 == 0, "vmcpde_unpack_upper_tiles: bad arguments");
  const int t = Pp / 128;
  pack_tiles_kernel<<<dim3(t, t), 256, 0, (cudaStream_t)stream>>>(nullptr, const_cast<double*>(packed), t, Pp, 1, S);
  VMC_LAUNCH_CHECK("pack_tiles_kernel(unpack)");
  return 0;
}

// In-place SUM all-reduce of `count` doubles on the caller's NCCL communicator (an ncclComm_t of the libnccl.so.2 loaded in
// this process) and stream.  The first-moment reduction of a right-hand side: [sum E, sum |E|, sum E^2, sum logp, sum O] (P + 4).
extern "C" __attribute__((visibility("default"))) int vmcpde_allreduce_sum(void* nccl_comm, double* buf, int64_t count, vmcpde_stream stream) {
  VMC_REQUIRE(nccl_comm && buf && count >= 0, "vmcpde_allreduce_sum: bad arguments");
  NcclApi* api = nullptr;
  if (int rc = nccl_api(&api)) return rc;
  const int rc = api->all_reduce(buf, buf, (size_t)count, /*ncclDouble*/ 8, /*ncclSum*/ 0, nccl_comm, (cudaStream_t)stream);
  if (rc != 0) return nccl_fail(api, "ncclAllReduce", rc);
  return 0;
}

// The second-moment reduction of a right-hand side: n_mats Pp x Pp Gram matrices (upper tiles valid) plus `n_tail` doubles
// (force vector and variance sums) are packed into `packed` (n_mats * vmcpde_packed_tiles_len(Pp) + n_tail doubles, caller
// owned), summed over ranks with ONE ncclAllReduce, and unpacked in place.  mats: HOST array of device pointers.
extern "C" __attribute__((visibility("default"))) int vmcpde_allreduce_moments(void* nccl_comm, double* const* mats, int32_t n_mats, int32_t Pp,
                                                                              double* tail, int64_t n_tail, double* packed,
                                                                              vmcpde_stream stream) {
  VMC_REQUIRE(nccl_comm && mats && packed && n_mats >= 0 && n_mats <= 8 && Pp > 0 && Pp % 128 == 0 && n_tail >= 0 && (tail || n_tail == 0),
              "vmcpde_allreduce_moments: bad arguments");
  const int64_t len = vmcpde_packed_tiles_len(Pp);
  cudaStream_t s = (cudaStream_t)stream;
  for (int m = 0; m < n_mats; ++m)
    if (int rc = vmcpde_pack_upper_tiles(mats[m], Pp, packed + m * len, stream)) return rc;
  if (n_tail) VMC_CUDA_CHECK(cudaMemcpyAsync(packed + n_mats * len, tail, (size_t)n_tail * 8, cudaMemcpyDeviceToDevice, s));
  if (int rc = vmcpde_allreduce_sum(nccl_comm, packed, n_mats * len + n_tail, stream)) return rc;
  for (int m = 0; m < n_mats; ++m)
    if (int rc = vmcpde_unpack_upper_tiles(packed + m * len, Pp, mats[m], stream)) return rc;
  if (n_tail) VMC_CUDA_CHECK(cudaMemcpyAsync(tail, packed + n_mats * len, (size_t)n_tail * 8, cudaMemcpyDeviceToDevice, s));
  return 0;
}

// Communicator plumbing for hosts without their own: id (128 bytes, host) from rank 0, shipped to the others out of band.
extern "C" __attribute__((visibility("default"))) int vmcpde_nccl_unique_id(char* id128) {
  VMC_REQUIRE(id128, "vmcpde_nccl_unique_id: null pointer");
  NcclApi* api = nullptr;
  if (int rc = nccl_api(&api)) return rc;
  NcclId id;
  const int rc = api->get_unique_id(&id);
  if (rc != 0) return nccl_fail(api, "ncclGetUniqueId", rc);
  std::memcpy(id128, id.internal, 128);
  return 0;
}

extern "C" __attribute__((visibility("default"))) int vmcpde_nccl_comm_init(int32_t n_ranks, int32_t rank, const char* id128, void** comm_out) {
  VMC_REQUIRE(id128 && comm_out && n_ranks >= 1 && rank >= 0 && rank < n_ranks, "vmcpde_nccl_comm_init: bad arguments");
  NcclApi* api = nullptr;
  if (int rc = nccl_api(&api)) return rc;
  NcclId id;
  std::memcpy(id.internal, id128, 128);
  const int rc = api->comm_init_rank(comm_out, n_ranks, id, rank);
  if (rc != 0) return nccl_fail(api, "ncclCommInitRank", rc);
  return 0;
}

extern "C" __attribute__((visibility("default"))) int vmcpde_nccl_comm_destroy(void* comm) {
  if (!comm) return 0;
  NcclApi* api = nullptr;
  if (int rc = nccl_api(&api)) return rc;
  const int rc = api->comm_destroy(comm);
  if (rc != 0) return nccl_fail(api, "ncclCommDestroy", rc);
  return 0;
}
