// bkm_rowpass.cu — label-indexed row passes for shapes whose k x d partial sums do not fit one CTA (the large-shape
// tensor path, bkm_tc2.cu: BASELINE config C5, k = 1024, d = 128 -> 512 KB of fp32 sums).
//
//   rowpass_mstep_kernel   M-step (_centers_dense, dask_ml/cluster/k_means.py:572-582 + da.bincount :548) from the
//                          labels the E-step wrote: the features are cut into DS slices of FS so that k x FS fp32 sums
//                          DO fit shared memory; CTA (row block rb, feature slice ds) sweeps its rows once, reading
//                          only its FS features of every row (whole 32-byte sectors: no HBM over-fetch), warp w owning
//                          the clusters c % NW == w (no atomics, fixed order).  Per-row-block partials are folded in
//                          float64 by reduce_partials in row-block order.
//   rowpass_dist_kernel    winning distance sum_i (x_i - c_label,i)^2 in direct form (fp32, no cancellation), one warp per
//                          row: min_out and the per-CTA distance sums (inertia, k_means.py:566 / k-means|| cost :466-469).
// The tensor-bound E-step dominates these shapes (C5: 22.6 ms at the tensor roof vs 5 ms for a sweep of X at the HBM
// roof), so the extra sweep costs ~1/5 of an iteration; shapes whose sums fit a CTA keep the fused kernels.
#include "bkm_common.cuh"
#include "bkm_ptx.cuh"
#include <cuda_bf16.h>
#include <math_constants.h>

namespace bkm {

static const int RP_THREADS = 512;
static const int RP_NW = RP_THREADS / 32;
static const int RP_TR = 256;                 // rows per staged tile

template <typename TX> __device__ __forceinline__ float rp_to_float(TX v);
template <> __device__ __forceinline__ float rp_to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float rp_to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

struct RpCfg {
  int FS, DS, RB;          // features per slice, feature slices, row blocks (grid = RB * DS)
  uint32_t off_sums, off_cnt, off_lab, off_x, total;
  uint32_t xrow_bytes;     // bytes of one staged row slice (FS * sizeof(TX))
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename TX>
__global__ void __launch_bounds__(RP_THREADS, 1)
rowpass_mstep_kernel(ChunkArgs a, RpCfg c) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int k = a.k, d = a.d, FS = c.FS;
  const int rb = blockIdx.x / c.DS, ds = blockIdx.x % c.DS;
  const int f0 = ds * FS;
  float* sums_s = reinterpret_cast<float*>(smem + c.off_sums);       // [k][FS]
  int* cnts_s = reinterpret_cast<int*>(smem + c.off_cnt);            // [k]
  int* lab_s = reinterpret_cast<int*>(smem + c.off_lab);             // [2][RP_TR]
  unsigned char* xs = smem + c.off_x;                                // [2][RP_TR][xrow_bytes]
  const TX* X = reinterpret_cast<const TX*>(a.X);
  for (int i = tid; i < k * FS; i += RP_THREADS) sums_s[i] = 0.f;
  for (int i = tid; i < k; i += RP_THREADS) cnts_s[i] = 0;

  const long long ntiles = (a.n + RP_TR - 1) / RP_TR;
  const int chunks = (int)(c.xrow_bytes / 16);                       // 16-byte chunks per row slice
  const uint32_t xs_u = ptx::smem_u32(xs), lab_u = ptx::smem_u32(lab_s);
  const uint32_t xbuf = (uint32_t)RP_TR * c.xrow_bytes;
  // stage tile t (labels + this CTA's feature slice of its rows) into buffer b; rows past n are labelled -1.
  // Chunks that start beyond the row pitch are skipped (their features are >= d and never read).
  auto stage = [&](long long t, int b) {
    const long long r0 = t * RP_TR;
    const int rows = (int)min((long long)RP_TR, a.n - r0);
    for (int e = tid; e < RP_TR * chunks; e += RP_THREADS) {
      const int r = e / chunks, q = e - r * chunks;
      const long long col = (long long)f0 + (long long)q * (16 / (int)sizeof(TX));
      if (r < rows && col < a.ldx)
        cp_async16(xs_u + (uint32_t)b * xbuf + (uint32_t)r * c.xrow_bytes + (uint32_t)q * 16u,
                   X + (r0 + r) * a.ldx + col);
    }
    for (int r = tid; r < RP_TR; r += RP_THREADS) {
      if (r < rows) cp_async4(lab_u + (uint32_t)(b * RP_TR + r) * 4u, a.labels + r0 + r);
      else lab_s[b * RP_TR + r] = -1;
    }
    cp_async_commit();
  };

  long long t = rb;
  int buf = 0;
  if (t < ntiles) stage(t, 0);
  __syncthreads();                                         // zeroed accumulators visible
#pragma unroll 1
  for (; t < ntiles; t += c.RB, buf ^= 1) {
    const long long tn = t + c.RB;
    if (tn < ntiles) { stage(tn, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    const int* lab = lab_s + buf * RP_TR;
    const unsigned char* xb = xs + (size_t)buf * xbuf;
    // warp w owns the clusters c with c % NW == w
#pragma unroll 1
    for (int base = 0; base < RP_TR; base += 32) {
      const int ml = lab[base + lane];
      unsigned m = __ballot_sync(0xffffffffu, ml >= 0 && (ml % RP_NW) == warp);
#pragma unroll 1
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        const int cl = __shfl_sync(0xffffffffu, ml, b);
        if (ds == 0 && lane == 0) cnts_s[cl] += 1;          // the warp owns cluster cl: no race
        const TX* xr = reinterpret_cast<const TX*>(xb + (size_t)(base + b) * c.xrow_bytes);
        float* sr = sums_s + (size_t)cl * FS;
        for (int f = lane; f < FS; f += 32)
          if (f0 + f < d) sr[f] += rp_to_float<TX>(xr[f]);
      }
    }
    __syncthreads();
  }
  // ---- flush this CTA's [k][FS] block into slot rb of the partial sums ----
  float* g = reinterpret_cast<float*>(a.psum) + (size_t)rb * k * d;
  for (int i = tid; i < k * FS; i += RP_THREADS) {
    const int cl = i / FS, f = i - cl * FS;
    if (f0 + f < d) g[(size_t)cl * d + f0 + f] = sums_s[i];
  }
  if (ds == 0) {
    int* gc = a.pcnt + (size_t)rb * k;
    for (int i = tid; i < k; i += RP_THREADS) gc[i] = cnts_s[i];
  }
}

template <typename TX>
__global__ void __launch_bounds__(256)
rowpass_dist_kernel(ChunkArgs a) {
  __shared__ double red_s[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int d = a.d, d4 = a.L.d4;
  const TX* X = reinterpret_cast<const TX*>(a.X);
  const float* gC = reinterpret_cast<const float*>(a.pack + a.L.off_cT);        // fp32 centres [k][d4]
  double dsum = 0.0;
  const long long gw = (long long)blockIdx.x * 8 + warp, nw = (long long)gridDim.x * 8;
  for (long long row = gw; row < a.n; row += nw) {
    const int lbl = a.labels[row];
    const TX* xr = X + row * a.ldx;
    const float* cr = gC + (size_t)(lbl < 0 ? 0 : lbl) * d4;
    float sacc = 0.f;
    for (int i = lane; i < d; i += 32) {
      const float df = rp_to_float<TX>(xr[i]) - cr[i];
      sacc = fmaf(df, df, sacc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
    if (lane == 0) {
      const float outv = a.squared ? sacc : sqrtf(sacc);
      dsum += (double)outv;
      if (a.min_out) reinterpret_cast<float*>(a.min_out)[row] = outv;
    }
  }
  if (lane == 0) red_s[warp] = dsum;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red_s[w];
    a.pin[blockIdx.x] = t;
  }
}

static bool make_rp_cfg(int k, int d, int esz, int sm_count, int psum_slots, long long n, RpCfg* c) {
  const Tc2Geom g = tc2_geom(k, d);
  c->FS = g.FS; c->DS = g.DS;
  int rb = sm_count / g.DS;
  if (rb < 1) rb = 1;
  if (rb > psum_slots) rb = psum_slots;
  const long long ntiles = (n + RP_TR - 1) / RP_TR;
  if (rb > ntiles) rb = (int)(ntiles > 0 ? ntiles : 1);
  c->RB = rb;
  c->xrow_bytes = (uint32_t)align_up((size_t)g.FS * esz, 16);
  uint32_t o = 0;
  c->off_sums = o; o += (uint32_t)k * g.FS * 4;
  c->off_cnt = o; o += (uint32_t)align_up((size_t)k * 4, 16);
  c->off_lab = o; o += 2 * RP_TR * 4;
  o = (uint32_t)align_up(o, 128);
  c->off_x = o; o += 2u * RP_TR * c->xrow_bytes;
  c->total = o;
  return o <= 227 * 1024;
}

int launch_rowpass_mstep(const ChunkArgs& a, int x_dtype, int sm_count, int* parts_out, cudaStream_t s) {
  if (!a.labels) return BKM_EINVAL;                     // the row pass is driven by the labels
  RpCfg c;
  const int esz = x_dtype == BKM_BF16 ? 2 : 4;
  if (!make_rp_cfg(a.k, a.d, esz, sm_count, a.psum_slots, a.n, &c)) return BKM_EUNSUPPORTED;
  if (x_dtype == BKM_BF16) {
    BKM_CUDA_TRY(cudaFuncSetAttribute(rowpass_mstep_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.total));
    rowpass_mstep_kernel<__nv_bfloat16><<<c.RB * c.DS, RP_THREADS, c.total, s>>>(a, c);
  } else {
    BKM_CUDA_TRY(cudaFuncSetAttribute(rowpass_mstep_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.total));
    rowpass_mstep_kernel<float><<<c.RB * c.DS, RP_THREADS, c.total, s>>>(a, c);
  }
  note_launch();
  BKM_CUDA_TRY(cudaGetLastError());
  *parts_out = c.RB;
  return 0;
}

int launch_rowpass_dist(const ChunkArgs& a, int x_dtype, int sm_count, int* parts_out, cudaStream_t s) {
  if (!a.labels) return BKM_EINVAL;
  long long grid = (a.n + 7) / 8;
  if (grid > (long long)sm_count * 8) grid = (long long)sm_count * 8;
  if (grid > a.part_slots) grid = a.part_slots;
  if (grid < 1) grid = 1;
  if (x_dtype == BKM_BF16) rowpass_dist_kernel<__nv_bfloat16><<<(int)grid, 256, 0, s>>>(a);
  else rowpass_dist_kernel<float><<<(int)grid, 256, 0, s>>>(a);
  note_launch();
  BKM_CUDA_TRY(cudaGetLastError());
  *parts_out = (int)grid;
  return 0;
}

}  // namespace bkm
