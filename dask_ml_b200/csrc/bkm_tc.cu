// bkm_tc.cu — fused E+M chunk kernel on the 5th-gen tensor cores (tcgen05 / TMEM / TMA), sm_100a.
//
// Shapes: fp32 X with d % 4 == 0, d <= 64, k <= 256 (BASELINE config C2: 10M x 64, k = 256).
//
// Per 128-row tile of X the kernel computes  acc = ||c||^2 + X . (-2 C)^T  with the product as a 3xTF32 split
//     Xhi.Bhi + Xhi.Blo + Xlo.Bhi          (B = -2C,  hi = tf32 part, lo = fp32 remainder)
// (~2^-21 relative error per product, fp32-GEMM class, instead of TF32's 2^-11) so that labels agree with the
// reference's float64 E-step (sklearn pairwise_distances_argmin_min, dask_ml/metrics/pairwise.py:35-38) except
// on near-ties, which are re-decided in float64; ||c||^2 enters through one extra K-step (rows [hi,mid,lo,0..]
// against a constant [1,1,1,0..] tile).  The M-step (_centers_dense, dask_ml/cluster/k_means.py:572-582) is
// fused: rows are scatter-added into REGISTER-resident per-CTA sums, X is read from HBM once.
//
// Warp roles (896 threads):
//   0       TMA producer of the A ring (2 x 128-row X tiles, SWIZZLE_128B K-major) + L2 prefetch ahead
//   1       MMA issuer (warp-converged, uniform-register operands, one elected lane issues tcgen05.mma/commit)
//   2       TMEM allocator (512 columns: 3 x 128 accumulator buffers + 2 x 64 Xlo)
//   3       TMA producer of the M ring (2 x 32-row quarter tiles re-fetched from L2 for the M-step warps)
//   4-7, 8-11   two Xlo-converter + epilogue warp sets on alternate tiles (thread == row == TMEM lane)
//   12-27   16 distance + M-step warps (warp w owns clusters c % 16 == w; lane l holds features l, l+32)
// Pipelines (mbarriers): A ring full/empty (empty is released by tcgen05.commit), accumulator full/empty per
// (epilogue set, buffer), Xlo full, labels full/empty, M ring full/empty.  DESIGN.md has the full description.
#include "bkm_common.cuh"
#include <cuda.h>
#include <math_constants.h>

namespace bkm {

static const int BM = 128;           // rows per tile
static const int NMW = 16;               // distance + M-step warps (each owns the clusters c % NMW == its index)
static const int TC_THREADS = (12 + NMW) * 32;
static const int KBLK_BYTES = BM * 128;   // one K-block (32 fp32 columns) of a 128-row tile
static const int MH = 32;                 // rows per M-ring stage (a quarter tile)
static const int MKBLK_BYTES = MH * 128;

struct TcCfg {
  int KB;        // 32-float K-blocks per row (1 or 2)
  int KS;        // MMA K-steps of 8 (ceil(d/8))
  int NP;        // padded centre count (multiple of 16, <= 256)
  int NU0, NU1;  // columns of unit 0 / unit 1 (NU1 == 0 -> one unit per tile)
  int U;
  int NST;       // X stages
  uint32_t off_bhi, off_blo, off_bcn, off_ones, off_x, off_m, off_cn, off_lab, off_flist, off_red, off_bar, off_tptr, total;
};

static const int NBUF = 3;               // 128-column TMEM accumulator buffers (3*128 + 2*64 Xlo = 512 columns)

enum {
  BAR_B_FULL = 0,
  BAR_X_FULL = 1,       // [NST<=4]
  BAR_X_EMPTY = 5,      // [4]
  BAR_ACC_FULL = 9,     // [set 2][buf 3]  one barrier per (epilogue warp set, accumulator buffer): every
  BAR_ACC_EMPTY = 15,   // [set 2][buf 3]  waiter then observes consecutive phases (no parity aliasing)
  BAR_XLO_FULL = 21,    // [2]
  BAR_LAB_FULL = 23,    // [2]
  BAR_LAB_EMPTY = 25,   // [2]
  BAR_M_FULL = 27,      // [2]  M ring: 64-row half tiles re-fetched (L2 hits) for the M-step warps
  BAR_M_EMPTY = 29,     // [2]
  BAR_COUNT = 31
};

// ------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Accumulator unit g (g = tile*U + u) uses TMEM buffer g % NBUF and is consumed by epilogue set (g/U) & 1.
// Parity of the phase of barrier [set][buf] that belongs to unit g = number of earlier units with the same
// (set, buffer), mod 2.  The pattern repeats every lcm(2U, 3) units: U=1 -> each pair once per 6 units,
// U=2 -> twice per 12 units (second occurrences: g % 12 in {4,6,8,9,10,11}).
__device__ __forceinline__ uint32_t acc_parity(long long g, int U) {
  if (U == 1) return (uint32_t)((g / 6) & 1);
  return (uint32_t)((0xF50u >> (int)(g % 12)) & 1u);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// A pipeline bug must fail, not hang the GPU: a wait that times out records (barrier offset, parity,
// warp) in g_tc_abort and returns; every later wait returns at once, the kernel drains with garbage and
// the host reports the code (bkm_debug_abort_code).
__device__ unsigned int g_tc_abort = 0;
__device__ unsigned int g_tc_dbg[64];     // per-warp abort code of the first CTA that aborts (debug)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"   // suspend-time hint: the hardware
        "selp.u32 %0, 1, 0, p;\n}"                                        // parks the warp instead of polling
        : "=r"(done) : "r"(bar), "r"(parity), "r"(2000u) : "memory");
    if (spin > (1u << 18) || ((spin & 1023) == 1023 && *(volatile unsigned int*)&g_tc_abort)) {
      atomicCAS(&g_tc_abort, 0u, 0x80000000u | ((bar & 0xfff) << 12) | (parity << 8) | (threadIdx.x >> 5));
      if (spin > 2048 && (threadIdx.x & 31) == 0) g_tc_dbg[threadIdx.x >> 5] = 0x80000000u | ((bar & 0xfff) << 12) | (parity << 8) | blockIdx.x;
      return;
    }
  }
}
// Long waits (a whole pipeline stage away): poll with back-off so that the poller does not steal
// issue slots from the warps doing the work on the same SM sub-partition.
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (!done) __nanosleep(128);
    if (spin > (1u << 20) || ((spin & 255) == 255 && *(volatile unsigned int*)&g_tc_abort)) {
      atomicCAS(&g_tc_abort, 0u, 0x80000000u | ((bar & 0xfff) << 12) | (parity << 8) | (threadIdx.x >> 5));
      return;
    }
  }
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* tm, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tm), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ float fmin3(float a, float b, float c) {
  float d;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));     // FMNMX3 on sm_100
  return d;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
#define TC_LD16(taddr, r)                                                                              \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                               \
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"                       \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),   \
                 "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]),            \
                 "=r"(r[13]), "=r"(r[14]), "=r"(r[15])                                                 \
               : "r"(taddr) : "memory")
#define TC_ST32(taddr, r)                                                                              \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                         \
               "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"                              \
               "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"                     \
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]),         \
                 "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]),       \
                 "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),   \
                 "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),   \
                 "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory")

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format, version 1):
// 8-row x 128-byte swizzle atoms, SBO = 1024 B between atoms along M/N, LBO unused (1).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                 // leading byte offset (>>4), ignored for swizzled K-major
  d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset (>>4)
  d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
  return d;
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = n
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
// byte offset of 16-byte chunk q (0..7) of row r inside one swizzled K-block
__device__ __forceinline__ uint32_t sw_chunk(int r, int q) { return (uint32_t)(r * 128 + ((q ^ (r & 7)) << 4)); }

#define ACC32_CASE(j) case j: acc[j][0] += x0; acc[j][1] += x1; break;

template <bool MSTEP, bool WANT_DIST>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_chunk_kernel(ChunkArgs a, TcCfg cfg, const __grid_constant__ CUtensorMap tm_x,
                const __grid_constant__ CUtensorMap tm_bhi, const __grid_constant__ CUtensorMap tm_blo,
                const __grid_constant__ CUtensorMap tm_xm) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t sbase = smem_u32(smem);
  if (tid == 0 && (sbase & 1023)) __trap();
  const uint32_t s_bhi = sbase + cfg.off_bhi, s_blo = sbase + cfg.off_blo, s_x = sbase + cfg.off_x;
  int* lab_s = reinterpret_cast<int*>(smem + cfg.off_lab);          // [2][BM]
  double* red_s = reinterpret_cast<double*>(smem + cfg.off_red);    // [5]
  const uint32_t bars = sbase + cfg.off_bar;
  uint32_t* tptr_s = reinterpret_cast<uint32_t*>(smem + cfg.off_tptr);
#define BAR(i) (bars + 8u * (uint32_t)(i))

  const int NST = cfg.NST, KB = cfg.KB, KS = cfg.KS, NP = cfg.NP, U = cfg.U;
  const uint32_t stage_bytes = (uint32_t)KB * KBLK_BYTES;
  const long long ntiles = (a.n + BM - 1) / BM;
  const long long my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  // ---------------- setup ----------------
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_x));
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_bhi));
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_blo));
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_xm));
    mbar_init(BAR(BAR_B_FULL), 1);
    for (int s = 0; s < 4; ++s) {
      mbar_init(BAR(BAR_X_FULL + s), 1);
      mbar_init(BAR(BAR_X_EMPTY + s), 1);
    }
    for (int b = 0; b < 2 * NBUF; ++b) {
      mbar_init(BAR(BAR_ACC_FULL + b), 1);
      mbar_init(BAR(BAR_ACC_EMPTY + b), 128);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(BAR(BAR_M_FULL + b), 1);
      mbar_init(BAR(BAR_M_EMPTY + b), NMW);
      mbar_init(BAR(BAR_XLO_FULL + b), 128);
      mbar_init(BAR(BAR_LAB_FULL + b), 128);
      mbar_init(BAR(BAR_LAB_EMPTY + b), NMW);
    }
    red_s[NMW] = 0.0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tptr_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  {
    // ||c||^2 enters the accumulator through one extra MMA K-step: B rows [hi,mid,lo,0,...] (exact 3-way
    // tf32 split, built by pack_norms_kernel) against a constant A tile of rows [1,1,1,0,...]; both tiles
    // use the canonical no-swizzle K-major layout (8-row groups of 256 B).  Plain stores + proxy fence.
    const float4* g = reinterpret_cast<const float4*>(a.pack + a.L.off_bcn);
    float4* sdst = reinterpret_cast<float4*>(smem + cfg.off_bcn);
    for (int i = tid; i < NP * 2; i += TC_THREADS) sdst[i] = g[i];
    float4* odst = reinterpret_cast<float4*>(smem + cfg.off_ones);
    for (int i = tid; i < BM * 2; i += TC_THREADS)
      odst[i] = ((i >> 3) & 1) ? make_float4(0.f, 0.f, 0.f, 0.f) : make_float4(1.f, 1.f, 1.f, 0.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tptr_s;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      mbar_expect_tx(BAR(BAR_B_FULL), 2u * (uint32_t)KB * (uint32_t)NP * 128u);
      for (int kb = 0; kb < KB; ++kb) {
        tma_load_2d(s_bhi + (uint32_t)kb * NP * 128u, &tm_bhi, BAR(BAR_B_FULL), kb * 32, 0);
        tma_load_2d(s_blo + (uint32_t)kb * NP * 128u, &tm_blo, BAR(BAR_B_FULL), kb * 32, 0);
      }
      const int PF = 6;        // L2 prefetch distance (tiles) ahead of the shared-memory ring
      for (long long it = 0; it < PF && it < my_tiles; ++it)
        for (int kb = 0; kb < KB; ++kb) tma_prefetch_2d(&tm_x, kb * 32, (int)((blockIdx.x + it * gridDim.x) * BM));
#pragma unroll 1
      for (long long it = 0; it < my_tiles; ++it) {
        const long long tile = blockIdx.x + it * gridDim.x;
        const int stage = (int)(it % NST);
        const uint32_t ph = (uint32_t)((it / NST) & 1);
        if (it + PF < my_tiles)
          for (int kb = 0; kb < KB; ++kb) tma_prefetch_2d(&tm_x, kb * 32, (int)((tile + (long long)PF * gridDim.x) * BM));
        mbar_wait(BAR(BAR_X_EMPTY + stage), ph ^ 1u);
        mbar_expect_tx(BAR(BAR_X_FULL + stage), stage_bytes);
        for (int kb = 0; kb < KB; ++kb)
          tma_load_2d(s_x + stage * stage_bytes + (uint32_t)kb * KBLK_BYTES, &tm_x, BAR(BAR_X_FULL + stage),
                      kb * 32, (int)(tile * BM));
      }
    }
  } else if (warp == 3) {
    // =========================== TMA producer of the M ring ===========================
    // Re-fetches every tile as two 64-row halves for the M-step / distance warps.  The same rows were
    // loaded for the MMA a few microseconds earlier, so these are L2 hits: HBM traffic stays at one read
    // of X per iteration while the M-step no longer holds the MMA's shared-memory stages.
    if (lane == 0) {
      const uint32_t s_m = sbase + cfg.off_m;
      const uint32_t mbytes = (uint32_t)KB * MKBLK_BYTES;
#pragma unroll 1
      for (long long mi = 0; mi < 4 * my_tiles; ++mi) {
        const long long tile = blockIdx.x + (mi >> 2) * gridDim.x;
        const int slot = (int)(mi & 1);
        mbar_wait(BAR(BAR_M_EMPTY + slot), (uint32_t)(((mi >> 1) & 1) ^ 1));
        mbar_expect_tx(BAR(BAR_M_FULL + slot), mbytes);
        for (int kb = 0; kb < KB; ++kb)
          tma_load_2d(s_m + slot * mbytes + (uint32_t)kb * MKBLK_BYTES, &tm_xm, BAR(BAR_M_FULL + slot), kb * 32,
                      (int)(tile * BM + (mi & 3) * MH));
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    // The whole warp runs this loop converged (all addresses / descriptors are warp-uniform and live
    // in uniform registers); only the tcgen05 instructions themselves are issued by one elected lane.
    const bool leader = elect_one();      // one lane issues every tcgen05.mma / commit (same thread: ordered)
    mbar_wait(BAR(BAR_B_FULL), 0);
    tc_fence_after();
    const uint64_t dflags = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    const uint32_t bblk = (uint32_t)NP * 128u;          // bytes of one K-block of a B tile
    // no-swizzle K-major operands of the ||c||^2 K-step: LBO = 128 B (second 16-byte K chunk), SBO = 256 B
    const uint64_t dns = ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46);
    const uint64_t dcn = dns | (uint64_t)(((sbase + cfg.off_bcn) >> 4) & 0x3FFF);
    const uint64_t dones = dns | (uint64_t)(((sbase + cfg.off_ones) >> 4) & 0x3FFF);
#pragma unroll 1
    for (long long it = 0; it < my_tiles; ++it) {
      const int stage = (int)(it % NST);
      const uint32_t ph = (uint32_t)((it / NST) & 1);
      const uint32_t xs = s_x + stage * stage_bytes;
      const uint32_t xlo_t = tmem + 384u + (uint32_t)(it & 1) * 64u;
      mbar_wait(BAR(BAR_X_FULL + stage), ph);
      tc_fence_after();
#pragma unroll 1
      for (int u = 0; u < U; ++u) {
        const long long g = it * U + u;
        const int buf = (int)(g % NBUF);
        if (g >= NBUF) {
          // the buffer was last used by unit g-NBUF, consumed by epilogue set ((g-NBUF)/U) & 1
          const long long gp = g - NBUF;
          mbar_wait(BAR(BAR_ACC_EMPTY + (int)((gp / U) & 1) * NBUF + buf), acc_parity(gp, U));
        }
        tc_fence_after();
        const int ncols = u == 0 ? cfg.NU0 : cfg.NU1;
        const uint32_t rowoff = u == 0 ? 0u : (uint32_t)cfg.NU0 * 128u;
        const uint32_t idesc = make_idesc(ncols);
        const uint32_t d_t = tmem + (uint32_t)buf * 128u;
        const uint64_t da = dflags | (uint64_t)((xs >> 4) & 0x3FFF);
        const uint64_t dbh = dflags | (uint64_t)(((s_bhi + rowoff) >> 4) & 0x3FFF);
        const uint64_t dbl = dflags | (uint64_t)(((s_blo + rowoff) >> 4) & 0x3FFF);
        // pass 1: Xhi . Bhi (A = raw fp32 tile; the tensor core reads its tf32 part), pass 2: Xhi . Blo
        if (leader) {
#pragma unroll
          for (int s = 0; s < 8; ++s) {
            if (s < KS) {
              const uint32_t ao = (uint32_t)(s >> 2) * (KBLK_BYTES >> 4) + (uint32_t)(s & 3) * 2u;
              const uint32_t bo = (uint32_t)(s >> 2) * (bblk >> 4) + (uint32_t)(s & 3) * 2u;
              mma_tf32_ss(d_t, da + ao, dbh + bo, idesc, s > 0 ? 1u : 0u);
            }
          }
#pragma unroll
          for (int s = 0; s < 8; ++s) {
            if (s < KS) {
              const uint32_t ao = (uint32_t)(s >> 2) * (KBLK_BYTES >> 4) + (uint32_t)(s & 3) * 2u;
              const uint32_t bo = (uint32_t)(s >> 2) * (bblk >> 4) + (uint32_t)(s & 3) * 2u;
              mma_tf32_ss(d_t, da + ao, dbl + bo, idesc, 1u);
            }
          }
        }
        __syncwarp();
        if (u == 0) {
          mbar_wait(BAR(BAR_XLO_FULL + (it & 1)), (uint32_t)((it >> 1) & 1));
          tc_fence_after();
        }
        // pass 3: Xlo . Bhi   (A from TMEM)
        if (leader) {
#pragma unroll
          for (int s = 0; s < 8; ++s) {
            if (s < KS) {
              const uint32_t bo = (uint32_t)(s >> 2) * (bblk >> 4) + (uint32_t)(s & 3) * 2u;
              mma_tf32_ts(d_t, xlo_t + (uint32_t)s * 8u, dbh + bo, idesc, 1u);
            }
          }
          mma_tf32_ss(d_t, dones, dcn + (uint64_t)(rowoff >> 6), idesc, 1u);     // + ||c_j||^2
          tc_commit(BAR(BAR_ACC_FULL + (int)(it & 1) * NBUF + buf));
          // the smem stage is free once every MMA of this tile has read it (the Xlo converter finished
          // before pass 3 could start); later consumers (M-step) read their rows from L2 instead
          if (u == U - 1) tc_commit(BAR(BAR_X_EMPTY + stage));
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // =========================== Xlo converter + epilogue ===========================
    // Two warp sets (warps 4-7 and 8-11) take alternate tiles, so each SM sub-partition has two epilogue
    // warps whose instruction streams interleave.  Thread == row == TMEM lane.
    const int set = (warp - 4) >> 2;
    const int q4 = warp & 3;
    const int r = q4 * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
    const float cnmax = (float)reinterpret_cast<const PackHeader*>(a.pack)->cn_max;
    uint32_t xoff[8];                              // swizzled 16-byte chunk offsets of this thread's row
#pragma unroll
    for (int q = 0; q < 8; ++q) xoff[q] = sw_chunk(r, q);
    mbar_wait(BAR(BAR_B_FULL), 0);
#pragma unroll 1
    for (long long it = set; it < my_tiles; it += 2) {
      const long long tile = blockIdx.x + it * gridDim.x;
      const int stage = (int)(it % NST);
      // ---- convert: Xlo = X - tf32(X) -> TMEM, and ||x||^2 ----
      float xn;
      {
        mbar_wait(BAR(BAR_X_FULL + stage), (uint32_t)((it / NST) & 1));
        const unsigned char* xs = smem + cfg.off_x + stage * stage_bytes;
        float xn4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
        for (int kb = 0; kb < KB; ++kb) {
          uint32_t v[32];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 t = *reinterpret_cast<const float4*>(xs + kb * KBLK_BYTES + xoff[q]);
            const float e[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              xn4[i] = fmaf(e[i], e[i], xn4[i]);
              const float hi = __uint_as_float(__float_as_uint(e[i]) & 0xFFFFE000u);
              v[q * 4 + i] = __float_as_uint(e[i] - hi);     // exact: the 13 low mantissa bits
            }
          }
          TC_ST32(tmem + lane_addr + 384u + (uint32_t)(it & 1) * 64u + (uint32_t)kb * 32u, v);
        }
        xn = (xn4[0] + xn4[1]) + (xn4[2] + xn4[3]);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        mbar_arrive(BAR(BAR_XLO_FULL + (it & 1)));
      }
      const float bound = a.tau * (xn + cnmax);
      // ---- two-pass epilogue per unit ----
      //  pass 1: m = min_j (S_j + ||c_j||^2)                      (1 FADD + 1/2 FMNMX3 per element)
      //  pass 2: every element within `bound` of m adds (1 + j/1024) to an accumulator on the FMA pipe:
      //          exactly one hit  -> acc = 1 + j/1024 : the arg-min, decoded exactly
      //          two or more hits -> acc >= 2        : near-tie, the row is deferred to float64
      float um0 = CUDART_INF_F, um1 = CUDART_INF_F, ua0 = 0.f, ua1 = 0.f;
#pragma unroll 1
      for (int u = 0; u < U; ++u) {
        const long long g = it * U + u;
        const int buf = (int)(g % NBUF);
        const int nch = (u == 0 ? cfg.NU0 : cfg.NU1) >> 4;
        const int col0 = u == 0 ? 0 : cfg.NU0;
        mbar_wait(BAR(BAR_ACC_FULL + set * NBUF + buf), acc_parity(g, U));
        tc_fence_after();
        const uint32_t tbase = tmem + lane_addr + (uint32_t)buf * 128u;
        uint32_t v0[16], v1[16];
        float ma = CUDART_INF_F, mb = CUDART_INF_F;
#define EPI_MIN(V, COLBASE)                                                              \
  {                                                                                      \
    _Pragma("unroll") for (int j4 = 0; j4 < 4; ++j4) {                                   \
      const float d0 = __uint_as_float(V[j4 * 4 + 0]);                                   \
      const float d1 = __uint_as_float(V[j4 * 4 + 1]);                                   \
      const float d2 = __uint_as_float(V[j4 * 4 + 2]);                                   \
      const float d3 = __uint_as_float(V[j4 * 4 + 3]);                                   \
      ma = fmin3(ma, d0, d1);                                                            \
      mb = fmin3(mb, d2, d3);                                                            \
    }                                                                                    \
  }
        TC_LD16(tbase, v0);
#pragma unroll 1
        for (int c = 0; c < nch; c += 2) {
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (c + 1 < nch) TC_LD16(tbase + (uint32_t)(c + 1) * 16u, v1);
          EPI_MIN(v0, col0 + c * 16)
          if (c + 1 < nch) {
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (c + 2 < nch) TC_LD16(tbase + (uint32_t)(c + 2) * 16u, v0);
            EPI_MIN(v1, col0 + (c + 1) * 16)
          }
        }
#undef EPI_MIN
        const float m = fminf(ma, mb);
        const float thr = m + bound;
        float acc = 0.f, accb = 0.f;
#define EPI_HIT(V, COLBASE)                                                              \
  {                                                                                      \
    float p0 = 0.f, p1 = 0.f;                                                            \
    _Pragma("unroll") for (int j4 = 0; j4 < 4; ++j4) {                                   \
      const float d0 = __uint_as_float(V[j4 * 4 + 0]);                                   \
      const float d1 = __uint_as_float(V[j4 * 4 + 1]);                                   \
      const float d2 = __uint_as_float(V[j4 * 4 + 2]);                                   \
      const float d3 = __uint_as_float(V[j4 * 4 + 3]);                                   \
      p0 = fmaf(d0 <= thr ? 1.f : 0.f, 1.f + (float)(j4 * 4 + 0) * 0.0009765625f, p0);   \
      p1 = fmaf(d1 <= thr ? 1.f : 0.f, 1.f + (float)(j4 * 4 + 1) * 0.0009765625f, p1);   \
      p0 = fmaf(d2 <= thr ? 1.f : 0.f, 1.f + (float)(j4 * 4 + 2) * 0.0009765625f, p0);   \
      p1 = fmaf(d3 <= thr ? 1.f : 0.f, 1.f + (float)(j4 * 4 + 3) * 0.0009765625f, p1);   \
    }                                                                                    \
    const float p = p0 + p1;                                                             \
    acc += p;                                                                            \
    accb += p >= 1.f ? (float)(COLBASE) : 0.f;                                           \
  }
        TC_LD16(tbase, v0);
#pragma unroll 1
        for (int c = 0; c < nch; c += 2) {
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (c + 1 < nch) TC_LD16(tbase + (uint32_t)(c + 1) * 16u, v1);
          EPI_HIT(v0, col0 + c * 16)
          if (c + 1 < nch) {
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (c + 2 < nch) TC_LD16(tbase + (uint32_t)(c + 2) * 16u, v0);
            EPI_HIT(v1, col0 + (c + 1) * 16)
          }
        }
#undef EPI_HIT
        tc_fence_before();
        mbar_arrive(BAR(BAR_ACC_EMPTY + set * NBUF + buf));
        // acc < 2: one hit, (acc - 1) * 1024 = index inside its 16-column chunk, accb = that chunk's base
        const float dec = acc < 2.f ? accb + (acc - 1.f) * 1024.f : -1.f;
        if (u == 0) { um0 = m; ua0 = dec; } else { um1 = m; ua1 = dec; }
      }
      // winner across units (unit 0 holds the lower indices: it wins exact ties)
      const bool win1 = um1 < um0;
      const float mw = win1 ? um1 : um0, mo = win1 ? um0 : um1, uaw = win1 ? ua1 : ua0;
      const long long row = tile * BM + r;
      const bool valid = row < a.n;
      const bool tie = uaw < 0.f || !(mo > mw + bound);
      const int bj = tie ? 0 : (int)(uaw + 0.5f);
      const bool flagged = valid && tie && a.k > 1;
      if (valid && !flagged && a.labels) a.labels[row] = bj;
      if (flagged) {
        // deferred: tc_recheck_kernel decides this row in float64 and adds its M-step contribution
        const int slot = atomicAdd(a.defer_cnt, 1);
        a.defer_idx[slot] = (int)row;
      }
      {
        const int lb = (int)(it & 1);
        mbar_wait(BAR(BAR_LAB_EMPTY + lb), (uint32_t)(((it >> 1) & 1) ^ 1));
        lab_s[lb * BM + r] = (valid && !flagged) ? bj : -1;
        mbar_arrive(BAR(BAR_LAB_FULL + lb));       // release semantics order the smem store
      }
    }
  } else if (warp >= 12) {
    // =========================== distance + M-step warps ===========================
    // Warp wm owns the rows whose label c satisfies c % 8 == wm.  Lane l holds features l and l+32 of
    // the row (conflict-free reads of the swizzled tile): (a) the winning distance is re-evaluated
    // exactly in fp32 direct form sum (x-c)^2 with c = -(bhi+blo)/2 read from the resident B tiles,
    // (b) [MSTEP] the row is added to the register-resident sums of cluster c.
    const int wm = warp - 12;
    float acc[256 / NMW][2];
#pragma unroll
    for (int j = 0; j < 256 / NMW; ++j) { acc[j][0] = 0.f; acc[j][1] = 0.f; }
    int cnt = 0;
    double inertia_acc = 0.0;
    const bool two = KB > 1;
    mbar_wait(BAR(BAR_B_FULL), 0);
#define MSTEP_ROW(XA, XB, CC, ROWG)                                                               \
  {                                                                                               \
    const int c = (CC);                                                                           \
    const float x0 = (XA), x1 = (XB);                                                             \
    if (WANT_DIST) {                                                                              \
      const uint32_t co = (uint32_t)(c * 128 + (((lane >> 2) ^ (c & 7)) << 4) + ((lane & 3) << 2)); \
      float t = fmaf(0.5f, *reinterpret_cast<const float*>(smem + cfg.off_bhi + co) +             \
                           *reinterpret_cast<const float*>(smem + cfg.off_blo + co), x0);         \
      float s2 = t * t;                                                                           \
      if (two) {                                                                                  \
        const uint32_t c1 = co + (uint32_t)NP * 128u;                                             \
        t = fmaf(0.5f, *reinterpret_cast<const float*>(smem + cfg.off_bhi + c1) +                 \
                       *reinterpret_cast<const float*>(smem + cfg.off_blo + c1), x1);             \
        s2 = fmaf(t, t, s2);                                                                      \
      }                                                                                           \
      _Pragma("unroll") for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o); \
      if (lane == 0) {                                                                            \
        const float outv = a.squared ? s2 : sqrtf(s2);                                            \
        inertia_acc += (double)outv;                                                              \
        if (a.min_out) reinterpret_cast<float*>(a.min_out)[(ROWG)] = outv;                        \
      }                                                                                           \
    }                                                                                             \
    if (MSTEP) {                                                                                  \
      const int cl = c / NMW;                                                                     \
      switch (cl) {                                                                               \
        ACC32_CASE(0) ACC32_CASE(1) ACC32_CASE(2) ACC32_CASE(3) ACC32_CASE(4) ACC32_CASE(5)       \
        ACC32_CASE(6) ACC32_CASE(7) ACC32_CASE(8) ACC32_CASE(9) ACC32_CASE(10) ACC32_CASE(11)     \
        ACC32_CASE(12) ACC32_CASE(13) ACC32_CASE(14) ACC32_CASE(15)                               \
        default: break;                                                                           \
      }                                                                                           \
      cnt += (lane == cl) ? 1 : 0;                                                                \
    }                                                                                             \
  }
#pragma unroll 1
    for (long long it = 0; it < my_tiles; ++it) {
      const long long tile = blockIdx.x + it * gridDim.x;
      const int lb = (int)(it & 1);
#pragma unroll 1
      for (int h = 0; h < 4; ++h) {
        const long long mi = it * 4 + h;
        const int slot = (int)(mi & 1);
        // every warp polls on its own: a per-quarter barrier across the 16 warps would make each quarter
        // as slow as its most loaded warp
        if (h == 0) mbar_wait(BAR(BAR_LAB_FULL + lb), (uint32_t)((it >> 1) & 1));
        mbar_wait(BAR(BAR_M_FULL + slot), (uint32_t)((mi >> 1) & 1));
        const unsigned char* xs = smem + cfg.off_m + slot * (KB * MKBLK_BYTES);
#pragma unroll 1
        for (int base = 0; base < MH; base += 32) {
          const int ml = lab_s[lb * BM + h * MH + base + lane];
          unsigned m = __ballot_sync(0xffffffffu, ml >= 0 && (ml & (NMW - 1)) == wm);
#pragma unroll 1
          while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            const int cq = __shfl_sync(0xffffffffu, ml, b) & 255;
            const int rh = base + b;      // row inside the 64-row half tile
            const uint32_t ro = (uint32_t)(rh * 128 + (((lane >> 2) ^ (rh & 7)) << 4) + ((lane & 3) << 2));
            const float xa = *reinterpret_cast<const float*>(xs + ro);
            const float xb = two ? *reinterpret_cast<const float*>(xs + MKBLK_BYTES + ro) : 0.f;
            MSTEP_ROW(xa, xb, cq, tile * BM + h * MH + base + b)
          }
        }
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(BAR(BAR_M_EMPTY + slot));
          if (h == 3) mbar_arrive(BAR(BAR_LAB_EMPTY + lb));
        }
      }
    }
#undef MSTEP_ROW
    if (lane == 0) red_s[wm] = inertia_acc;
    if (MSTEP) {
      // flush the register-resident sums: cluster c = wm + 8 j, features lane and lane + 32
      float* g = reinterpret_cast<float*>(a.psum) + (size_t)blockIdx.x * a.k * a.d;
#pragma unroll
      for (int j = 0; j < 256 / NMW; ++j) {
        const int c = wm + NMW * j;
        if (c < a.k) {
          if (lane < a.d) g[(size_t)c * a.d + lane] = acc[j][0];
          if (lane + 32 < a.d) g[(size_t)c * a.d + lane + 32] = acc[j][1];
        }
      }
      const int cc = wm + NMW * lane;
      if (lane < 256 / NMW && cc < a.k) a.pcnt[(size_t)blockIdx.x * a.k + cc] = cnt;
    }
  }

  // ---------------- teardown ----------------
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    double t = red_s[NMW];
    for (int w = 0; w < NMW; ++w) t += red_s[w];
    a.pin[blockIdx.x] = t;
  }
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
#undef BAR
}

// ------------------------------------------------------------------------------------------
// Deferred float64 re-check.  Rows whose 3xTF32 best/second margin was inside the rounding bound
// were left out of the fused kernel's outputs and M-step; here they are decided exactly:
// d2_j = sum_i (x_i - c_ji)^2 in float64 against the float64 centres (transposed in shared memory so
// that thread j <-> centre j reads are conflict-free), lowest index on exact ties.  Their labels,
// distances and M-step contributions are then added (float64 atomics: order-insensitive to ~1e-16).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tc_recheck_kernel(ChunkArgs a, bool mstep, double* sums, unsigned long long* counts, double* dist_sum) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int cnt = *a.defer_cnt;
  if ((int)blockIdx.x >= cnt) return;
  const int k = a.k, d = a.d, tid = threadIdx.x;
  const int kpad = (k + 31) / 32 * 32 + 1;                 // odd pitch: conflict-free transpose writes
  double* cT = reinterpret_cast<double*>(smem);             // [d][kpad]
  float* xrow = reinterpret_cast<float*>(smem + (size_t)d * kpad * 8);   // [d]
  double* wd = reinterpret_cast<double*>(xrow + ((d + 3) & ~3));         // [8]
  int* wj = reinterpret_cast<int*>(wd + 8);                               // [8]
  const double* gC64 = reinterpret_cast<const double*>(a.pack + a.L.off_c64);
  for (int e = tid; e < k * d; e += 256) { const int j = e / d, i = e - j * d; cT[i * kpad + j] = gC64[e]; }
  const float* X = reinterpret_cast<const float*>(a.X);
  for (int f = blockIdx.x; f < cnt; f += gridDim.x) {
    const long long row = a.defer_idx[f];
    __syncthreads();
    for (int i = tid; i < d; i += 256) xrow[i] = X[row * a.ldx + i];
    __syncthreads();
    double bd = CUDART_INF; int bj = 0x7fffffff;
    for (int j = tid; j < k; j += 256) {
      double s0 = 0.0, s1 = 0.0;
      int i = 0;
      for (; i + 1 < d; i += 2) {
        const double d0 = (double)xrow[i] - cT[i * kpad + j];
        const double d1 = (double)xrow[i + 1] - cT[(i + 1) * kpad + j];
        s0 = fma(d0, d0, s0); s1 = fma(d1, d1, s1);
      }
      if (i < d) { const double d0 = (double)xrow[i] - cT[i * kpad + j]; s0 = fma(d0, d0, s0); }
      const double sdist = s0 + s1;
      if (sdist < bd) { bd = sdist; bj = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double od = __shfl_xor_sync(0xffffffffu, bd, o);
      const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
      if (od < bd || (od == bd && oj < bj)) { bd = od; bj = oj; }
    }
    if ((tid & 31) == 0) { wd[tid >> 5] = bd; wj[tid >> 5] = bj; }
    __syncthreads();
    double fd = wd[0]; int fj = wj[0];
    for (int w = 1; w < 8; ++w) if (wd[w] < fd || (wd[w] == fd && wj[w] < fj)) { fd = wd[w]; fj = wj[w]; }
    if (tid == 0) {
      const double outv = a.squared ? fd : sqrt(fd);
      if (a.labels) a.labels[row] = fj;
      if (a.min_out) reinterpret_cast<float*>(a.min_out)[row] = (float)outv;
      if (dist_sum) atomicAdd(dist_sum, outv);
      if (mstep) atomicAdd(counts + fj, 1ull);
    }
    if (mstep) for (int i = tid; i < d; i += 256) atomicAdd(sums + (size_t)fj * d + i, (double)xrow[i]);
  }
}

static int launch_tc_recheck(const ChunkArgs& a, bool mstep, int sm_count, cudaStream_t s) {
  const int kpad = (a.k + 31) / 32 * 32 + 1;
  const size_t smem = (size_t)a.d * kpad * 8 + (size_t)((a.d + 3) & ~3) * 4 + 8 * 8 + 8 * 4 + 16;
  if (smem > 227 * 1024) return BKM_EUNSUPPORTED;
  BKM_CUDA_TRY(cudaFuncSetAttribute(tc_recheck_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_recheck_kernel<<<sm_count, 256, smem, s>>>(a, mstep, a.out_sums, (unsigned long long*)a.out_counts, a.out_dist_sum);
  note_launch();
  BKM_CUDA_TRY(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D fp32 tensor [rows][cols] with row pitch `pitch_elems`, box = 32 columns x box_rows, 128B swizzle
static int make_map(CUtensorMap* tm, const void* base, long long rows, int cols, long long pitch_elems, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return BKM_EUNSUPPORTED;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)pitch_elems * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : BKM_EUNSUPPORTED;
}

unsigned int tc_abort_code() {
  unsigned int v = 0;
  cudaMemcpyFromSymbol(&v, g_tc_abort, sizeof(v));
  return v;
}
void tc_abort_detail(unsigned int* out64) { cudaMemcpyFromSymbol(out64, g_tc_dbg, 64 * sizeof(unsigned int)); }

bool tc_supported(int d, int k, int dtype) {
  return dtype == BKM_F32 && d >= 4 && d <= 64 && (d % 4) == 0 && k >= 1 && k <= 256;
}

static bool make_cfg(int d, int k, TcCfg* c) {
  c->KB = (d + 31) / 32;
  c->KS = (d + 7) / 8;
  c->NP = (k + 15) / 16 * 16;
  if (c->NP <= 128) { c->NU0 = c->NP; c->NU1 = 0; c->U = 1; }
  else { c->NU0 = (c->NP / 2 + 15) / 16 * 16; c->NU1 = c->NP - c->NU0; c->U = c->NU1 > 0 ? 2 : 1; }
  const uint32_t bbytes = (uint32_t)c->KB * c->NP * 128u;
  for (int nst = 3; nst >= 2; --nst) {
    uint32_t o = 0;
    c->off_bhi = o; o += bbytes;
    c->off_blo = o; o += bbytes;
    o = (uint32_t)align_up(o, 1024);
    c->off_bcn = o; o += (uint32_t)c->NP * 32u;                  // ||c||^2 operand tile
    c->off_ones = o; o += BM * 32u;                              // constant [1,1,1,0..] A tile
    o = (uint32_t)align_up(o, 1024);
    c->off_x = o; o += (uint32_t)nst * c->KB * KBLK_BYTES;       // A ring (MMA + Xlo converter)
    c->off_m = o; o += 2u * c->KB * MKBLK_BYTES;                 // M ring (M-step / distance warps)
    c->off_cn = o;
    c->off_lab = o; o += 2 * BM * 4;
    c->off_flist = o;
    c->off_red = o; o += (NMW + 1) * 8;
    c->off_bar = o; o += BAR_COUNT * 8;
    c->off_tptr = o; o += 16;
    c->total = o;
    c->NST = nst;
    if (o <= 227 * 1024) return true;
  }
  return false;
}

int launch_tc(const ChunkArgs& a, bool mstep, int sm_count, int* grid_out, cudaStream_t s) {
  if ((reinterpret_cast<uintptr_t>(a.X) & 15) || (a.ldx % 4)) return BKM_EALIGN;
  TcCfg cfg;
  if (!make_cfg(a.d, a.k, &cfg)) return BKM_EUNSUPPORTED;
  CUtensorMap tm_x, tm_bhi, tm_blo, tm_xm;
  int rc = make_map(&tm_x, a.X, a.n, a.d, a.ldx, BM);
  if (rc) return rc;
  rc = make_map(&tm_xm, a.X, a.n, a.d, a.ldx, MH);
  if (rc) return rc;
  rc = make_map(&tm_bhi, a.pack + a.L.off_bhi, a.L.kp, a.L.dk, a.L.dk, cfg.NP);
  if (rc) return rc;
  rc = make_map(&tm_blo, a.pack + a.L.off_blo, a.L.kp, a.L.dk, a.L.dk, cfg.NP);
  if (rc) return rc;
  long long ntiles = (a.n + BM - 1) / BM;
  int grid = (int)(ntiles < sm_count ? ntiles : sm_count);
  if (grid < 1) grid = 1;
  *grid_out = grid;
  const bool want_dist = !mstep || a.min_out != nullptr || a.want_sum;
  BKM_CUDA_TRY(cudaMemsetAsync(a.defer_cnt, 0, sizeof(int), s));
#define TC_LAUNCH(M, W)                                                                                       \
  {                                                                                                           \
    BKM_CUDA_TRY(cudaFuncSetAttribute(tc_chunk_kernel<M, W>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                      (int)cfg.total));                                                       \
    tc_chunk_kernel<M, W><<<grid, TC_THREADS, cfg.total, s>>>(a, cfg, tm_x, tm_bhi, tm_blo, tm_xm);                  \
  }
  if (mstep) { if (want_dist) TC_LAUNCH(true, true) else TC_LAUNCH(true, false) }
  else TC_LAUNCH(false, true)
#undef TC_LAUNCH
  note_launch(2);
  BKM_CUDA_TRY(cudaGetLastError());
  if (a.k > 1) {
    int rc2 = launch_tc_recheck(a, mstep, sm_count, s);
    if (rc2) return rc2;
  }
  note_launch();
  BKM_CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace bkm
