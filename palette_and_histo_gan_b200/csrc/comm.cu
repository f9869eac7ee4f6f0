// The one exchange of the sharded path as a kernel over peer memory (NVLink / NVSwitch P2P).
//
// The Hellinger distance takes one square root over the whole batch (histogram.py:88-89), so ranks that each hold
// a shard of the batch must add up one double per step (the sum of squares; optionally a second value).  A library
// all-reduce of 8 bytes costs ~30 us of launch + protocol latency — 3 % of a 1 ms step at 8 GPUs — and, through
// torch.distributed, a handful of framework kernels around it.  Here it is ONE warp: every rank stores its value
// into a mailbox slot in the HBM of every peer (plain P2P stores, ordered by a release at system scope), then polls
// its OWN mailbox (local memory) until the values of all ranks for this round have arrived, and adds them in rank
// order — every rank gets bit-identical sums.  Two slot sets alternate by round: a rank can run at most one round
// ahead of a peer (it needs the peer's value of round e+1 to finish round e+1, and the peer publishes that only
// after it finished round e), so a slot is never overwritten while somebody still reads it.
#include <string.h>

#include <new>

#include "common.cuh"
#include "hist_internal.cuh"

namespace ph {

struct CommSlot {
  double v[2];
  unsigned long long epoch;
  unsigned long long pad;
};
static_assert(sizeof(CommSlot) == 32, "mailbox slot");

struct CommParams {
  CommSlot* peer[PH_COMM_MAX_WORLD];  // mailbox of every rank as mapped into this process (own = local)
  CommSlot* local;
  unsigned long long epoch;
  int rank, world;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void __launch_bounds__(32) comm_allreduce_f64_kernel(CommParams P, double* value, int count) {
  const int lane = threadIdx.x;
  const int parity = (int)(P.epoch & 1ull);
  pdl_wait();  // the value was produced by the kernel before this one in the stream
  pdl_launch_dependents();
  double a0 = 0.0, a1 = 0.0;
  if (lane < P.world) {
    // publish: this rank's value into slot [parity][rank] of rank `lane`
    CommSlot* dst = P.peer[lane] + parity * PH_COMM_MAX_WORLD + P.rank;
    const double v0 = value[0], v1 = count > 1 ? value[1] : 0.0;
    *reinterpret_cast<volatile double*>(&dst->v[0]) = v0;
    *reinterpret_cast<volatile double*>(&dst->v[1]) = v1;
    __threadfence_system();
    st_release_sys(&dst->epoch, P.epoch);
    // gather: rank `lane`'s value from the local mailbox
    const CommSlot* src = P.local + parity * PH_COMM_MAX_WORLD + lane;
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(&src->epoch) != P.epoch) {
      __nanosleep(200);
      if (global_ns() - t0 > 10000000000ull) {  // a peer that never arrives must not hang the GPU
        printf("palhist: rank %d waited 10 s for rank %d in round %llu of the peer all-reduce\n", P.rank, lane, P.epoch);
        __trap();
      }
    }
    a0 = *reinterpret_cast<const volatile double*>(&src->v[0]);
    a1 = *reinterpret_cast<const volatile double*>(&src->v[1]);
  }
  __syncwarp();
  double s0 = 0.0, s1 = 0.0;
  for (int r = 0; r < P.world; ++r) {  // rank order: the same rounding on every rank
    s0 += __shfl_sync(0xffffffffu, a0, r);
    s1 += __shfl_sync(0xffffffffu, a1, r);
  }
  if (lane == 0) {
    value[0] = s0;
    if (count > 1) value[1] = s1;
  }
}

}  // namespace ph

struct ph_comm {
  int device = 0, rank = 0, world = 1;
  ph::CommSlot* local = nullptr;
  ph::CommSlot* peer[PH_COMM_MAX_WORLD] = {};
  bool opened[PH_COMM_MAX_WORLD] = {};
  unsigned long long epoch = 0;
  bool connected = false;
};

using namespace ph;

extern "C" {

int ph_comm_create(int device, int rank, int world, ph_comm** out) {
  PH_CHECK_ARG(out != nullptr, "comm output pointer is NULL");
  PH_CHECK_ARG(world >= 1 && world <= PH_COMM_MAX_WORLD && rank >= 0 && rank < world, "bad rank %d / world %d (at most %d ranks)",
               rank, world, PH_COMM_MAX_WORLD);
  PH_CUDA_OK(cudaSetDevice(device));
  ph_comm* c = new (std::nothrow) ph_comm();
  PH_CHECK_ARG(c != nullptr, "out of host memory");
  c->device = device; c->rank = rank; c->world = world;
  const size_t bytes = 2 * PH_COMM_MAX_WORLD * sizeof(CommSlot);
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&c->local), bytes);
  if (e == cudaSuccess) e = cudaMemset(c->local, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    set_error("mailbox allocation failed: %s", cudaGetErrorString(e));
    if (c->local) cudaFree(c->local);
    delete c;
    return PH_ERR_CUDA;
  }
  c->peer[rank] = c->local;
  c->connected = world == 1;
  *out = c;
  return PH_OK;
}

int ph_comm_export(ph_comm* comm, void* handle_host) {
  PH_CHECK_ARG(comm && handle_host, "NULL pointer argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == PH_COMM_HANDLE_BYTES, "IPC handle size");
  PH_CUDA_OK(cudaSetDevice(comm->device));
  cudaIpcMemHandle_t h;
  PH_CUDA_OK(cudaIpcGetMemHandle(&h, comm->local));
  memcpy(handle_host, &h, sizeof(h));
  return PH_OK;
}

int ph_comm_connect(ph_comm* comm, const void* handles_host) {
  PH_CHECK_ARG(comm && handles_host, "NULL pointer argument");
  PH_CUDA_OK(cudaSetDevice(comm->device));
  for (int r = 0; r < comm->world; ++r) {
    if (r == comm->rank || comm->opened[r]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const char*>(handles_host) + (size_t)r * PH_COMM_HANDLE_BYTES, sizeof(h));
    void* p = nullptr;
    PH_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    comm->peer[r] = static_cast<CommSlot*>(p);
    comm->opened[r] = true;
  }
  comm->connected = true;
  return PH_OK;
}

int ph_comm_allreduce_sum_f64(ph_comm* comm, double* value, int count, void* stream) {
  PH_CHECK_ARG(comm && value, "NULL pointer argument");
  PH_CHECK_ARG(count == 1 || count == 2, "count must be 1 or 2");
  PH_CHECK_ARG(comm->connected, "ph_comm_allreduce_sum_f64 before ph_comm_connect");
  CommParams P{};
  for (int r = 0; r < comm->world; ++r) P.peer[r] = comm->peer[r];
  P.local = comm->local;
  P.epoch = ++comm->epoch;
  P.rank = comm->rank;
  P.world = comm->world;
  PH_CUDA_OK(launch_pdl(comm_allreduce_f64_kernel, dim3(1), dim3(32), 0, static_cast<cudaStream_t>(stream), P, value, count));
  PH_LAUNCH_OK("comm_allreduce_f64_kernel");
  return PH_OK;
}

void ph_comm_destroy(ph_comm* comm) {
  if (!comm) return;
  cudaSetDevice(comm->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < comm->world; ++r)
    if (comm->opened[r]) cudaIpcCloseMemHandle(comm->peer[r]);
  if (comm->local) cudaFree(comm->local);
  delete comm;
}

}  // extern "C"
