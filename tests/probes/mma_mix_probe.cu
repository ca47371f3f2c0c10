// mma_mix_probe.cu — hardware probe (not a test of the product): does tcgen05.mma kind::f16 accept DIFFERENT 16-bit
// formats for A and B (bf16 x fp16)?  The instruction descriptor has independent a_format / b_format fields; the
// large-shape tensor kernel (bkm_tc2.cu) multiplies raw bf16 rows of X by fp16 (hi, lo) splits of the centres when this
// works, and falls back to bf16 splits otherwise.  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gpurun_out/mma_mix_probe tests/probes/mma_mix_probe.cu && gpurun_out/mma_mix_probe
// Prints one line per (A format, B format) pair: max |D - reference| over a 128 x 32 tile, K = 16.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// A: 128 x 16 (16-bit), B: 32 x 16 (16-bit), both K-major in the canonical NO-swizzle layout:
// 8-row x 16-byte core matrices; the two K chunks of a row group are 128 B apart (LBO), row groups 256 B apart (SBO).
__global__ void probe(const uint16_t* A, const uint16_t* B, float* D, int afmt, int bfmt) {
  __shared__ __align__(128) uint16_t sa[128 * 16];
  __shared__ __align__(128) uint16_t sb[32 * 16];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tptr;
  const int tid = threadIdx.x;
  for (int i = tid; i < 128 * 16; i += blockDim.x) {
    const int r = i / 16, kk = i % 16;
    sa[(r / 8) * 128 + (kk / 8) * 64 + (r % 8) * 8 + (kk % 8)] = A[i];
  }
  for (int i = tid; i < 32 * 16; i += blockDim.x) {
    const int r = i / 16, kk = i % 16;
    sb[(r / 8) * 128 + (kk / 8) * 64 + (r % 8) * 8 + (kk % 8)] = B[i];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tptr;
  if (tid == 0) {
    const uint64_t dns = ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46);
    const uint64_t da = dns | (uint64_t)((smem_u32(sa) >> 4) & 0x3FFF);
    const uint64_t db = dns | (uint64_t)((smem_u32(sb) >> 4) & 0x3FFF);
    // fp32 accumulate, K-major A and B, M = 128, N = 32
    const uint32_t idesc = (1u << 4) | ((uint32_t)afmt << 7) | ((uint32_t)bfmt << 10) | ((uint32_t)(32 >> 3) << 17) |
                           ((uint32_t)(128 >> 4) << 24);
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
        ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(0u) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // everyone waits for the MMA
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid < 128) {
    const int warp = tid >> 5;
    uint32_t v[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) D[tid * 32 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32));
}

static uint16_t enc(float v, int fmt) {
  if (fmt == 1) { __nv_bfloat16 b = __float2bfloat16(v); return *reinterpret_cast<uint16_t*>(&b); }
  __half h = __float2half(v);
  return *reinterpret_cast<uint16_t*>(&h);
}
static float dec(uint16_t u, int fmt) {
  if (fmt == 1) { __nv_bfloat16 b = *reinterpret_cast<__nv_bfloat16*>(&u); return __bfloat162float(b); }
  __half h = *reinterpret_cast<__half*>(&u);
  return __half2float(h);
}

int main() {
  uint16_t hA[128 * 16], hB[32 * 16];
  float hD[128 * 32];
  uint16_t *dA, *dB;
  float* dD;
  cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD, sizeof(hD));
  const char* names[2] = {"f16", "bf16"};
  int rc = 0;
  for (int afmt = 0; afmt < 2; ++afmt)
    for (int bfmt = 0; bfmt < 2; ++bfmt) {
      srand(1234);
      // values that are exact in one format only (many mantissa bits / wide range) so that a mis-decoded operand shows
      for (int i = 0; i < 128 * 16; ++i) hA[i] = enc((float)(rand() % 2001 - 1000) / 7.0f * (afmt == 1 ? 64.0f : 1.0f), afmt);
      for (int i = 0; i < 32 * 16; ++i) hB[i] = enc((float)(rand() % 2001 - 1000) / 3.0f, bfmt);
      cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
      cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
      cudaMemset(dD, 0, sizeof(hD));
      probe<<<1, 128>>>(dA, dB, dD, afmt, bfmt);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("A=%s B=%s : CUDA error %s\n", names[afmt], names[bfmt], cudaGetErrorString(e)); rc = 1; break; }
      cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
      double worst = 0, mag = 0;
      for (int r = 0; r < 128; ++r)
        for (int j = 0; j < 32; ++j) {
          double s = 0;
          for (int kk = 0; kk < 16; ++kk) s += (double)dec(hA[r * 16 + kk], afmt) * (double)dec(hB[j * 16 + kk], bfmt);
          worst = fmax(worst, fabs(s - (double)hD[r * 32 + j]));
          mag = fmax(mag, fabs(s));
        }
      printf("A=%s B=%s : max |D - ref| = %.6g (max |ref| = %.6g, relative %.3g) %s\n", names[afmt], names[bfmt], worst, mag,
             worst / mag, worst / mag < 1e-5 ? "OK" : "MISMATCH");
    }
  return rc;
}
