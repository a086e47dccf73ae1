// Stand-alone validation (run on the GPU box) of the 16-bit operand scheme of the v6 propagate engine:
//   tile X (128 x 128 fp32 in [0,1])  ->  two fp16 planes  h0 = fp16(x), h1 = fp16((x - h0) * 2^11)   (22-bit mantissa)
//   features F (128 x 32 fp32)        ->  three bf16 pieces f0 + f1 + f2                             (24-bit mantissa)
//   one SWIZZLE_NONE shared-memory image per plane, element (i, j) at (j/8)*SJ + (i/8)*128 + (i%8)*16 + (j%8)*2, read
//     as a K-major  A operand (M = i, K = j) for   D1 = X   * F      and
//     as an MN-major A operand (M = j, K = i) for  D2 = X^T * F      -- no transposition pass at all;
//   tcgen05.mma kind::f16 with A = f16, B = bf16 (mixed), K = 16 per instruction:
//     D[:, 0:96]  += h0 * [f0|f1|f2]        D[:, 96:160] += h1 * [f0|f1]          y = c0+c1+c2 + 2^-11 (c3+c4)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include "../mc-gra_b200/csrc/tc_common.cuh"

constexpr int T = 128, KC = 32;
constexpr uint32_t SJ = 16 * 128 + 16;        // stride between 8-column groups of a plane
constexpr uint32_t LBO_B = 96 * 16 + 16;      // stride between 8-k groups of the B block (96 rows of 16 B)

__device__ __forceinline__ uint32_t make_idesc_f16(int M, int N, int a_bf16, int b_bf16, int a_mn, int b_mn) {
  uint32_t d = 0;
  d |= 1u << 4;                                // D = F32
  d |= (uint32_t)(a_bf16 & 1) << 7;            // A format: 0 f16, 1 bf16
  d |= (uint32_t)(b_bf16 & 1) << 10;
  d |= (uint32_t)(a_mn & 1) << 15;
  d |= (uint32_t)(b_mn & 1) << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
               "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(128) k_test(const float* Xg, const float* Fg, float* D1g, float* D2g, int swap_mn, int which_mma, int a_bf16) {
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned char* p0 = sm;                      // fp16 plane h0
  unsigned char* p1 = sm + 16 * SJ;            // fp16 plane h1
  unsigned char* bb = sm + 32 * SJ;            // bf16 B block: rows n = [f0 (32) | f1 (32) | f2 (32)], K-major
  __shared__ uint32_t tmem_base;
  __shared__ uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < T * T; e += 128) {
    const int i = e / T, j = e % T;
    const float x = Xg[e];
    const uint32_t off = (uint32_t)(j >> 3) * SJ + (uint32_t)(i >> 3) * 128u + (uint32_t)(i & 7) * 16u + (uint32_t)(j & 7) * 2u;
    if (a_bf16) {      // (format probe only: bf16 planes lose precision)
      const __nv_bfloat16 b0 = __float2bfloat16_rn(x);
      *reinterpret_cast<__nv_bfloat16*>(p0 + off) = b0;
      *reinterpret_cast<__nv_bfloat16*>(p1 + off) = __float2bfloat16_rn((x - __bfloat162float(b0)) * 2048.f);
    } else {
      const __half h0 = __float2half_rn(x);
      *reinterpret_cast<__half*>(p0 + off) = h0;
      *reinterpret_cast<__half*>(p1 + off) = __float2half_rn((x - __half2float(h0)) * 2048.f);
    }
  }
  for (int e = tid; e < T * KC; e += 128) {
    const int k = e / KC, c = e % KC;
    const float f = Fg[e];
    const __nv_bfloat16 f0 = __float2bfloat16_rn(f);
    const float r1 = f - __bfloat162float(f0);
    const __nv_bfloat16 f1 = __float2bfloat16_rn(r1);
    const __nv_bfloat16 f2 = __float2bfloat16_rn(r1 - __bfloat162float(f1));
    const __nv_bfloat16 v[3] = {f0, f1, f2};
    for (int p = 0; p < 3; ++p) {
      const int n = p * KC + c;
      *reinterpret_cast<__nv_bfloat16*>(bb + (uint32_t)(k >> 3) * LBO_B + (uint32_t)(n >> 3) * 128u + (uint32_t)(n & 7) * 16u + (uint32_t)(k & 7) * 2u) = v[p];
    }
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base, 512);
  if (tid == 0) tc::mbar_init(&bar, 1);
  tc::fence_async_smem();
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tm = tmem_base;
  if (tid == 0) {
    const uint32_t id96 = make_idesc_f16(128, 96, a_bf16, 1, 0, 0), id64 = make_idesc_f16(128, 64, a_bf16, 1, 0, 0);
    const uint32_t id96t = make_idesc_f16(128, 96, a_bf16, 1, 1, 0), id64t = make_idesc_f16(128, 64, a_bf16, 1, 1, 0);
    for (int ks = 0; ks < T / 16; ++ks) {
      const uint64_t bd = tc::make_desc(tc::smem_u32(bb) + ks * 2 * LBO_B, LBO_B, 128u);
      // direct: A K-major (M = i: SBO = 128 between 8-row groups; K = j: LBO = SJ between 8-column groups)
      if (which_mma & 1) {
      mma_f16(tm + 0, tc::make_desc(tc::smem_u32(p0) + ks * 2 * SJ, SJ, 128u), bd, id96, ks > 0);
      mma_f16(tm + 96, tc::make_desc(tc::smem_u32(p1) + ks * 2 * SJ, SJ, 128u), bd, id64, ks > 0);
      }
      if (!(which_mma & 2)) continue;
      // mirrored: A MN-major (M = j, K = i) on the same image; K step = 16 rows i = 2 groups of 128 B
      const uint32_t lbo = swap_mn ? SJ : 128u, sbo = swap_mn ? 128u : SJ;
      mma_f16(tm + 160, tc::make_desc(tc::smem_u32(p0) + ks * 2 * 128, lbo, sbo), bd, id96t, ks > 0);
      mma_f16(tm + 256, tc::make_desc(tc::smem_u32(p1) + ks * 2 * 128, lbo, sbo), bd, id64t, ks > 0);
    }
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::fence_after();
  const uint32_t tl = tm + ((uint32_t)(warp * 32) << 16);
  for (int which = 0; which < 2; ++which) {
    float c[5][32];
    for (int p = 0; p < 5; ++p) tc::tmem_ld32(tl + (which ? 160 : 0) + p * 32, c[p]);
    float* dst = (which ? D2g : D1g) + (warp * 32 + lane) * KC;
    for (int k = 0; k < KC; ++k) dst[k] = (c[0][k] + c[1][k] + c[2][k]) + (c[3][k] + c[4][k]) * (1.f / 2048.f);
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tm, 512);
}

int main(int argc, char** argv) {
  const int which_mma = argc > 1 ? atoi(argv[1]) : 3, a_bf16 = argc > 2 ? atoi(argv[2]) : 0;
  std::vector<float> X(T * T), F(T * KC), D1(T * KC), D2(T * KC);
  srand(7);
  for (auto& v : X) { const float u = (float)rand() / RAND_MAX; v = (rand() % 4 == 0) ? u * 1e-4f : u; }
  for (auto& v : F) v = ((float)rand() / RAND_MAX - 0.5f) * ((rand() % 8 == 0) ? 1e-6f : 3.f);
  float *dX, *dF, *dD1, *dD2;
  cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dF, F.size() * 4); cudaMalloc(&dD1, D1.size() * 4); cudaMalloc(&dD2, D2.size() * 4);
  cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dF, F.data(), F.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = 32 * SJ + 16 * LBO_B + 1024;
  cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int swap_mn = 0; swap_mn < 2; ++swap_mn) {
    cudaMemset(dD1, 0, D1.size() * 4); cudaMemset(dD2, 0, D2.size() * 4);
    k_test<<<1, 128, smem>>>(dX, dF, dD1, dD2, swap_mn, which_mma, a_bf16);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D1.data(), dD1, D1.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(D2.data(), dD2, D2.size() * 4, cudaMemcpyDeviceToHost);
    double e1 = 0, e2 = 0, ref = 0;
    for (int m = 0; m < T; ++m)
      for (int c = 0; c < KC; ++c) {
        double s1 = 0, s2 = 0, a1 = 0;
        for (int k = 0; k < T; ++k) { s1 += (double)X[m * T + k] * F[k * KC + c]; s2 += (double)X[k * T + m] * F[k * KC + c]; a1 += fabs((double)X[m * T + k] * F[k * KC + c]); }
        e1 = fmax(e1, fabs(D1[m * KC + c] - s1) / a1);
        e2 = fmax(e2, fabs(D2[m * KC + c] - s2) / a1);
        ref = fmax(ref, fabs(s1));
      }
    printf("mma set %d, A %s | MN-major desc %s: direct max err / sum|terms| = %.3e   mirrored = %.3e   (ref max %.3f)  D1[0..1]=%.5f %.5f D2[0..1]=%.5f %.5f\n",
           which_mma, a_bf16 ? "bf16" : "f16", swap_mn ? "(LBO=SJ, SBO=128)" : "(LBO=128, SBO=SJ)", e1, e2, ref, D1[0], D1[1], D2[0], D2[1]);
  }
  return 0;
}
