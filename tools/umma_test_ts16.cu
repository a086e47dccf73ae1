// Stand-alone validation (GPU box) of kind::f16 tcgen05.mma with the A operand in TENSOR MEMORY:
//   A (128 x 128 fp16, two planes h0 / h1 of an fp32 tile in [0,1]) written by 128 threads with tcgen05.st, thread = lane =
//   M index, two consecutive K elements packed per 32-bit column (order = argv[1]: 0 low half = even k, 1 = odd k);
//   B = fp16 [g0 | g1] block, K-major SWIZZLE_NONE (the layout of propagate's k_prep_b16);  D = A * B for both the tile
//   (lane = row) and its transpose (lane = column).  Prints max error / sum |terms| against fp64.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include "../mc-gra_b200/csrc/tc_common.cuh"

constexpr int T = 128, KC = 32;
constexpr uint32_t LBO_B = (2 * KC) * 16;      // stride between 8-k groups of the B block (2KC rows of 16 B)

__device__ __forceinline__ uint32_t idesc(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
               "r"(a_tmem), "l"(bdesc), "r"(id), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(128) k_test(const float* Xg, const float* Fg, float* D1g, float* D2g, int order) {
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned char* bb = sm;
  __shared__ uint32_t tmem_base;
  __shared__ uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < T * KC; e += 128) {
    const int k = e / KC, c = e % KC;
    const float g = Fg[e];
    const __half g0 = __float2half_rn(g);
    const __half g1 = __float2half_rn((g - __half2float(g0)) * 2048.f);
    unsigned char* blk = bb + (uint32_t)(k >> 3) * LBO_B + (uint32_t)(k & 7) * 2u;
    *reinterpret_cast<__half*>(blk + (uint32_t)(c >> 3) * 128u + (uint32_t)(c & 7) * 16u) = g0;
    const int c2 = c + KC;
    *reinterpret_cast<__half*>(blk + (uint32_t)(c2 >> 3) * 128u + (uint32_t)(c2 & 7) * 16u) = g1;
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base, 512);
  if (tid == 0) tc::mbar_init(&bar, 1);
  tc::fence_async_smem();
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tm = tmem_base;
  // images: columns [128, 192) T h0 | [192, 256) T h1 | [256, 320) T^T h0 | [320, 384) T^T h1;  D1 [0, 96), D2 [384, 480)
  const uint32_t tl = tm + ((uint32_t)(warp * 32) << 16);
  for (int img = 0; img < 2; ++img) {
    for (int c16 = 0; c16 < 4; ++c16) {        // 16 columns = 32 K elements at a time
      uint32_t h[16], l[16];
      for (int u = 0; u < 16; ++u) {
        float v[2];
        for (int w = 0; w < 2; ++w) {
          const int k = c16 * 32 + u * 2 + w;
          v[w] = img == 0 ? Xg[tid * T + k] : Xg[k * T + tid];
        }
        const __half a0 = __float2half_rn(v[0]), a1 = __float2half_rn(v[1]);
        const __half r0 = __float2half_rn((v[0] - __half2float(a0)) * 2048.f), r1 = __float2half_rn((v[1] - __half2float(a1)) * 2048.f);
        const uint32_t ua0 = __half_as_ushort(a0), ua1 = __half_as_ushort(a1), ur0 = __half_as_ushort(r0), ur1 = __half_as_ushort(r1);
        h[u] = order ? (ua1 | (ua0 << 16)) : (ua0 | (ua1 << 16));
        l[u] = order ? (ur1 | (ur0 << 16)) : (ur0 | (ur1 << 16));
      }
      tc::tmem_st16(tl + 128 + img * 128 + c16 * 16, h);
      tc::tmem_st16(tl + 192 + img * 128 + c16 * 16, l);
    }
  }
  tc::tmem_st_wait();
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  if (tid == 0) {
    const uint32_t id_cat = idesc(128, 2 * KC), id_one = idesc(128, KC);
    for (int ks = 0; ks < T / 16; ++ks) {
      const uint64_t bd = tc::make_desc(tc::smem_u32(bb) + ks * 2 * LBO_B, LBO_B, 128u);
      mma_f16_ts(tm + 0, tm + 128 + ks * 8, bd, id_cat, ks > 0);          // T h0 x [g0 | g1]
      mma_f16_ts(tm + 64, tm + 192 + ks * 8, bd, id_one, ks > 0);         // T h1 x g0
      mma_f16_ts(tm + 384, tm + 256 + ks * 8, bd, id_cat, ks > 0);        // T^T h0 x [g0 | g1]
      mma_f16_ts(tm + 448, tm + 320 + ks * 8, bd, id_one, ks > 0);        // T^T h1 x g0
    }
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::fence_after();
  for (int which = 0; which < 2; ++which) {
    float c[3][32];
    for (int p = 0; p < 3; ++p) tc::tmem_ld32(tl + (which ? 384 : 0) + p * 32, c[p]);
    float* dst = (which ? D2g : D1g) + (warp * 32 + lane) * KC;
    for (int k = 0; k < KC; ++k) dst[k] = c[0][k] + (c[1][k] + c[2][k]) * (1.f / 2048.f);
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tm, 512);
}

int main(int argc, char** argv) {
  std::vector<float> X(T * T), F(T * KC), D1(T * KC), D2(T * KC);
  srand(7);
  for (auto& v : X) { const float u = (float)rand() / RAND_MAX; v = (rand() % 4 == 0) ? u * 1e-4f : u; }
  for (auto& v : F) v = ((float)rand() / RAND_MAX - 0.5f) * ((rand() % 8 == 0) ? 1e-3f : 3.f);
  float *dX, *dF, *dD1, *dD2;
  cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dF, F.size() * 4); cudaMalloc(&dD1, D1.size() * 4); cudaMalloc(&dD2, D2.size() * 4);
  cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dF, F.data(), F.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = 16 * LBO_B + 1024;
  cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int order = 0; order < 2; ++order) {
    cudaMemset(dD1, 0, D1.size() * 4); cudaMemset(dD2, 0, D2.size() * 4);
    k_test<<<1, 128, smem>>>(dX, dF, dD1, dD2, order);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D1.data(), dD1, D1.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(D2.data(), dD2, D2.size() * 4, cudaMemcpyDeviceToHost);
    double e1 = 0, e2 = 0;
    for (int m = 0; m < T; ++m)
      for (int c = 0; c < KC; ++c) {
        double s1 = 0, s2 = 0, a1 = 0, a2 = 0;
        for (int k = 0; k < T; ++k) {
          s1 += (double)X[m * T + k] * F[k * KC + c]; a1 += fabs((double)X[m * T + k] * F[k * KC + c]);
          s2 += (double)X[k * T + m] * F[k * KC + c]; a2 += fabs((double)X[k * T + m] * F[k * KC + c]);
        }
        e1 = fmax(e1, fabs(D1[m * KC + c] - s1) / a1);
        e2 = fmax(e2, fabs(D2[m * KC + c] - s2) / a2);
      }
    printf("pack order %d (low half = %s k): direct max err / sum|terms| = %.3e   mirrored = %.3e\n", order, order ? "odd" : "even", e1, e2);
  }
  return 0;
}
