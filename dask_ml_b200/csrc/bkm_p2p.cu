// bkm_p2p.cu — the per-iteration collective of the Lloyd loop as ONE kernel over NVLink peer memory.
//
// dask_ml/cluster/k_means.py:545-550 folds the per-chunk partials of all workers (`da.atop(..., sum)` + bincount); with one
// process per GPU that is a sum all-reduce of [k*d sums | k counts | inertia] (133 KB at 10M x 64, k = 256) per iteration.
// At that size the NCCL call is pure latency (20-28 us measured + launch), so every rank instead PUSHES its buffer into a
// mailbox slot on every peer (plain stores through the NVSwitch), raises a flag there, waits for the flags in its own
// mailbox and adds the slots in RANK ORDER — the same order on every rank, so all ranks hold bit-identical sums (what the
// stop test of the device-resident loop needs) and the result is reproducible run to run.
//
// Mailbox of rank q (one cudaMalloc allocation, exported with cudaIpcGetMemHandle, opened by every peer):
//   [parity 2][world] uint32 flags | [2] uint32 arrival counters | [parity 2][world][max_elems] float64 slots
// Call number `seq` uses parity seq & 1: a slot of parity p is written again by call seq + 2, which a peer can only
// start after it has seen THIS rank's flag of call seq + 1, which this rank raises after it has finished reading call seq.
#include "bkm_common.cuh"
#include "bkm_ptx.cuh"
#include <math_constants.h>
#include <string.h>
#include <stdlib.h>

namespace bkm {

static const int P2P_MAX_WORLD = 64;
static const size_t P2P_FLAGS_OFF = 0;        // [2][64] uint32
static const size_t P2P_CNT_OFF = 512;        // [2] uint32
static const size_t P2P_DATA_OFF = 1024;

__host__ __device__ inline size_t p2p_slot_off(int parity, int src, int world, long long max_elems) {
  return P2P_DATA_OFF + ((size_t)parity * world + src) * (size_t)max_elems * 8;
}

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(256)
p2p_allreduce_kernel(double* __restrict__ buf, long long n, unsigned char* const* __restrict__ boxes, int rank, int world,
                     long long max_elems, unsigned seq) {
  __shared__ bool last_s;
  __shared__ int fail_s;
  const int tid = threadIdx.x;
  const int par = (int)(seq & 1u);
  unsigned char* mine = boxes[rank];
  const long long gt = (long long)blockIdx.x * blockDim.x + tid, nth = (long long)gridDim.x * blockDim.x;
  // ---- A: push this rank's buffer into slot [par][rank] of every mailbox (its own included) ----
  for (int q = 0; q < world; ++q) {
    double* dst = reinterpret_cast<double*>(boxes[q] + p2p_slot_off(par, rank, world, max_elems));
    for (long long i = gt; i < n; i += nth) dst[i] = buf[i];
  }
  __threadfence_system();
  __syncthreads();
  if (tid == 0) {
    unsigned* cnt = reinterpret_cast<unsigned*>(mine + P2P_CNT_OFF) + par;
    last_s = atomicAdd(cnt, 1u) == gridDim.x - 1;
    fail_s = 0;
  }
  __syncthreads();
  if (last_s) {
    // every CTA of this rank has pushed (and fenced): raise this rank's flag on every peer
    __threadfence_system();
    if (tid < world) st_release_sys(reinterpret_cast<unsigned*>(boxes[tid] + P2P_FLAGS_OFF) + par * P2P_MAX_WORLD + rank, seq);
    if (tid == 0) reinterpret_cast<unsigned*>(mine + P2P_CNT_OFF)[par] = 0u;      // for call seq + 2
  }
  // ---- B: wait for every rank's flag in this rank's mailbox (wall-clock limit: a dead peer must not hang the GPU) ----
  if (tid < world) {
    const unsigned* f = reinterpret_cast<const unsigned*>(mine + P2P_FLAGS_OFF) + par * P2P_MAX_WORLD + tid;
    const unsigned long long t0 = ptx::globaltimer_ns();
    unsigned spin = 0;
    while (ld_acquire_sys(f) != seq) {
      if ((++spin & 1023u) == 0u && ptx::globaltimer_ns() - t0 > 2000000000ull) { atomicExch(&fail_s, 1); break; }
    }
  }
  __syncthreads();
  // ---- C: add the slots in rank order (L1 is bypassed: the slots were written by other GPUs) ----
  if (fail_s) {
    if (gt == 0) buf[0] = CUDART_NAN;          // loud: the host turns a non-finite shift into an error
    return;
  }
  for (long long i = gt; i < n; i += nth) {
    double s = 0.0;
    for (int src = 0; src < world; ++src)
      s += __ldcg(reinterpret_cast<const double*>(mine + p2p_slot_off(par, src, world, max_elems)) + i);
    buf[i] = s;
  }
}

}  // namespace bkm

using namespace bkm;

extern "C" {

int bkm_p2p_mailbox_bytes(int world, int64_t max_elems, size_t* nbytes) {
  if (!nbytes || world < 1 || world > P2P_MAX_WORLD || max_elems < 1) return BKM_EINVAL;
  *nbytes = p2p_slot_off(2, 0, world, max_elems);
  return 0;
}

int bkm_p2p_alloc(size_t nbytes, void** dev_ptr) {
  if (!dev_ptr || nbytes == 0) return BKM_EINVAL;
  BKM_CUDA_TRY(cudaMalloc(dev_ptr, nbytes));         // a whole cudaMalloc allocation: that is what an IPC handle names
  BKM_CUDA_TRY(cudaMemset(*dev_ptr, 0, nbytes));
  BKM_CUDA_TRY(cudaDeviceSynchronize());
  return 0;
}

int bkm_p2p_free(void* dev_ptr) {
  if (dev_ptr) BKM_CUDA_TRY(cudaFree(dev_ptr));
  return 0;
}

int bkm_p2p_export(void* dev_ptr, void* handle64_host) {
  if (!dev_ptr || !handle64_host) return BKM_EINVAL;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  BKM_CUDA_TRY(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64_host), dev_ptr));
  return 0;
}

int bkm_p2p_import(const void* handle64_host, void** dev_ptr) {
  if (!handle64_host || !dev_ptr) return BKM_EINVAL;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64_host, sizeof(h));
  BKM_CUDA_TRY(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int bkm_p2p_close(void* dev_ptr) {
  if (dev_ptr) BKM_CUDA_TRY(cudaIpcCloseMemHandle(dev_ptr));
  return 0;
}

int bkm_allreduce_p2p(double* buf, int64_t n, void* const* mailboxes_dev, int rank, int world, int64_t max_elems,
                      unsigned int seq, void* stream) {
  if (!buf || !mailboxes_dev || n < 0 || n > max_elems || world < 1 || world > P2P_MAX_WORLD || rank < 0 || rank >= world)
    return BKM_EINVAL;
  if (n == 0) return 0;
  static const int grid_cap = [] { const char* e = getenv("BKM_P2P_GRID"); int v = e ? atoi(e) : 0; return v >= 1 && v <= 32 ? v : 32; }();
  long long grid = (n + 1023) / 1024;               // <= 32 CTAs: all resident at once (they wait for one another)
  if (grid > grid_cap) grid = grid_cap;
  if (grid < 1) grid = 1;
  p2p_allreduce_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(buf, n, reinterpret_cast<unsigned char* const*>(mailboxes_dev),
                                                                    rank, world, max_elems, seq);
  note_launch();
  BKM_CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // extern "C"
