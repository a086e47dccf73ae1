// Stand-alone validation of the tcgen05 building blocks used by the tensor-core engines (run on the GPU box):
//   * SWIZZLE_NONE K-major and MN-major tf32 operands written by CUDA threads, descriptors from tc_common.cuh
//   * D[128 x 64] in TMEM, read back with tcgen05.ld
//   * how kind::tf32 treats the low 13 mantissa bits of fp32 inputs (truncate vs round)
//   * the 3xTF32 composition  A_hi*[B_hi|B_lo] (N=64) + A_lo*B_hi (N=32)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cstring>
#include "../mc-gra_b200/csrc/tc_common.cuh"

constexpr int M = 128, K = 128, N = 64;
constexpr uint32_t SJ = 2048 + 16;   // stride between 4-element groups of the "j" index (padded: conflict-free stores)
constexpr uint32_t SI = 128;         // stride between 8-row groups of the "i" index

// tile layout shared by both orientations: element (i, j) at (j/4)*SJ + (i/8)*SI + (i%8)*16 + (j%4)*4
__device__ __forceinline__ uint32_t tile_off(int i, int j) { return (j >> 2) * SJ + (i >> 3) * SI + (i & 7) * 16 + (j & 3) * 4; }
// MN-major tf32 A operand in the SW128_32B layout (the only MN-major tf32 layout on sm_100): element (k=i, mn=j):
//   (j/32)*LBO_MN + (i/4)*512 + (i%4)*128 + ((((j%32)/8) ^ (i%4))*32) + (j%8)*4     [32 B chunks XOR-swizzled by k row]
constexpr uint32_t LBO_MN = 32 * 512;
__device__ __forceinline__ uint32_t mn_off(int i, int j) {
  return (uint32_t)(j >> 5) * LBO_MN + (uint32_t)(i >> 2) * 512u + (uint32_t)(i & 3) * 128u +
         (uint32_t)((((j & 31) >> 3) ^ (i & 3)) * 32) + (uint32_t)(j & 7) * 4u;
}
// B operand, MN-major (n contiguous): element (k, n) at (n/4)*SBO_B + (k/8)*128 + (k%8)*16 + (n%4)*4
constexpr uint32_t SBO_B = 16 * 128;
__device__ __forceinline__ uint32_t b_off(int k, int n) { return (n >> 2) * SBO_B + (k >> 3) * 128 + (k & 7) * 16 + (n & 3) * 4; }

// B operand, K-major (k contiguous): element (k, n) at (k/4)*LBO_BK + (n/8)*128 + (n%8)*16 + (k%4)*4
constexpr uint32_t LBO_BK = 8 * 128 + 16;
__device__ __forceinline__ uint32_t bk_off(int k, int n) { return (k >> 2) * LBO_BK + (n >> 3) * 128 + (n & 7) * 16 + (k & 3) * 4; }

__global__ void __launch_bounds__(128) k_stld(float* out, uint32_t* info) {
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tc::tmem_alloc(&tmem_base, 32);
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tm = tmem_base;
  if (tid == 0) info[0] = tm;
  const uint32_t addr = tm + ((uint32_t)(warp * 32) << 16);
  for (int c = 0; c < 32; ++c) {
    uint32_t v = __float_as_uint((float)(tid * 100 + c));
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(addr + c), "r"(v) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  float v[32];
  tc::tmem_ld32(addr, v);
  for (int q = 0; q < 32; ++q) out[tid * 32 + q] = v[q];
  tc::fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tm, 32);
}

__global__ void __launch_bounds__(128) k_test(const float* Ag, const float* Bg, float* Dg, int mode, int split, int bmode) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* a_hi = smem;                       // 32 * SJ
  unsigned char* a_lo = a_hi + 32 * SJ;
  unsigned char* b_sm = a_lo + 32 * SJ;             // N=64 (hi | lo): 16 * SBO_B
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < M * K; e += 128) {
    const int i = e / K, j = e % K;
    float v = Ag[e];
    float hi = split ? __uint_as_float(__float_as_uint(v) & 0xffffe000u) : v;
    if (mode == 2) {
      *reinterpret_cast<float*>(a_hi + mn_off(i, j)) = hi;
      *reinterpret_cast<float*>(a_lo + mn_off(i, j)) = v - hi;
    } else {
      *reinterpret_cast<float*>(a_hi + tile_off(i, j)) = hi;
      *reinterpret_cast<float*>(a_lo + tile_off(i, j)) = v - hi;
    }
  }
  for (int e = tid; e < K * 32; e += 128) {
    const int k = e / 32, n = e % 32;
    float v = Bg[k * 32 + n];
    float hi = split ? __uint_as_float(__float_as_uint(v) & 0xffffe000u) : v;
    if (bmode == 1) {
      *reinterpret_cast<float*>(b_sm + b_off(k, n)) = hi;
      *reinterpret_cast<float*>(b_sm + b_off(k, n + 32)) = split ? (v - hi) : 0.f;
    } else {
      *reinterpret_cast<float*>(b_sm + bk_off(k, n)) = hi;
      *reinterpret_cast<float*>(b_sm + bk_off(k, n + 32)) = split ? (v - hi) : 0.f;
    }
  }
  if (tid == 0) tc::mbar_init(&bar, 1);
  tc::fence_async_smem();
  __syncthreads();
  if (warp == 0) tc::tmem_alloc(&tmem_base, mode == 3 ? 512 : 64);
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tm = tmem_base;
  if (mode == 3) {     // A operand into TMEM: lane m = tid, column k: hi at [64, 192), lo at [192, 320)
    const uint32_t base = tm + ((uint32_t)(warp * 32) << 16);
    for (int k0 = 0; k0 < K; k0 += 16) {
      uint32_t h[16], l[16];
      for (int q = 0; q < 16; ++q) {
        const float v = Ag[(k0 + q) * K + tid];          // A[m][k] = Ag[k][m]
        const float hi = split ? __uint_as_float(__float_as_uint(v) & 0xffffe000u) : v;
        h[q] = __float_as_uint(hi); l[q] = __float_as_uint(v - hi);
      }
      tc::tmem_st16(base + 64 + k0, h);
      tc::tmem_st16(base + 192 + k0, l);
    }
    tc::tmem_st_wait();
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
  }
  if (tid == 0) {
    const uint32_t id64 = tc::make_idesc_tf32(128, 64, mode >= 1, bmode);
    const uint32_t id32 = tc::make_idesc_tf32(128, 32, mode >= 1, bmode);
    for (int ks = 0; ks < K / 8; ++ks) {
      uint64_t ad_hi, ad_lo;
      if (mode == 0) {        // A[m=i][k=j]: K-major, SBO = SI (8-row groups of i), LBO = SJ (4-groups of j)
        ad_hi = tc::make_desc(tc::smem_u32(a_hi) + ks * 2 * SJ, SJ, SI);
        ad_lo = tc::make_desc(tc::smem_u32(a_lo) + ks * 2 * SJ, SJ, SI);
      } else if (mode == 2) { // A[m=j][k=i]: MN-major SW128_32B: LBO = stride between 32-wide mn atoms, SBO = between 4-row k atoms
        ad_hi = tc::make_desc_sw(tc::smem_u32(a_hi) + ks * 2 * 512, LBO_MN, 512, 1);
        ad_lo = tc::make_desc_sw(tc::smem_u32(a_lo) + ks * 2 * 512, LBO_MN, 512, 1);
      } else {                // A[m=j][k=i]: MN-major, SBO = SJ (4-groups of j = mn), LBO = SI (8-groups of i = k)
        ad_hi = tc::make_desc(tc::smem_u32(a_hi) + ks * SI, SI, SJ);
        ad_lo = tc::make_desc(tc::smem_u32(a_lo) + ks * SI, SI, SJ);
      }
      if (mode == 3) {
        const uint64_t bd3 = tc::make_desc(tc::smem_u32(b_sm) + ks * 2 * LBO_BK, LBO_BK, 128);
        tc::mma_tf32_ts(tm, tm + 64 + ks * 8, bd3, tc::make_idesc_tf32(128, 64, 0, 0), ks > 0);
        if (split) tc::mma_tf32_ts(tm, tm + 192 + ks * 8, bd3, tc::make_idesc_tf32(128, 32, 0, 0), 1);
        continue;
      }
      const uint64_t bd = bmode == 1 ? tc::make_desc(tc::smem_u32(b_sm) + ks * 128, 128, SBO_B)
                                     : tc::make_desc(tc::smem_u32(b_sm) + ks * 2 * LBO_BK, LBO_BK, 128);
      tc::mma_tf32(tm, ad_hi, bd, id64, ks > 0);
      if (split) tc::mma_tf32(tm, ad_lo, bd, id32, 1);
    }
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::fence_after();
  float v[32];
  for (int c = 0; c < 2; ++c) {
    tc::tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c * 32, v);
    for (int q = 0; q < 32; ++q) Dg[(warp * 32 + (tid & 31)) * 64 + c * 32 + q] = v[q];
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tm, mode == 3 ? 512 : 64);
}

static float trunc_tf32(float v) { uint32_t u; memcpy(&u, &v, 4); u &= 0xffffe000u; memcpy(&v, &u, 4); return v; }
static float rna_tf32(float v) { uint32_t u; memcpy(&u, &v, 4); u += 0x1000u; u &= 0xffffe000u; memcpy(&v, &u, 4); return v; }

int main() {
  std::vector<float> A(M * K), B(K * 32), D(M * 64);
  srand(7);
  for (auto& x : A) x = (float)rand() / RAND_MAX;                 // like the adjacency estimate: [0, 1]
  for (auto& x : B) x = ((float)rand() / RAND_MAX - 0.5f) * 4.f;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = 2 * 32 * SJ + 32 * LBO_BK + 16 * SBO_B + 1024;
  cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  {
    float* dO; uint32_t* dI; cudaMalloc(&dO, 128 * 32 * 4); cudaMalloc(&dI, 16);
    k_stld<<<1, 128>>>(dO, dI);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> O(128 * 32); uint32_t info[4];
    cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(info, dI, 16, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int t = 0; t < 128; ++t) for (int c = 0; c < 32; ++c) bad += (O[t * 32 + c] != (float)(t * 100 + c));
    printf("st/ld roundtrip: err=%s tmem_base=0x%08x mismatches=%d  sample O[5*32+3]=%.1f\n", cudaGetErrorString(e), info[0], bad, O[5 * 32 + 3]);
  }
  for (int bmode = 0; bmode < 1; ++bmode)
  for (int mode = 0; mode < 4; ++mode)
    for (int split = 0; split < 2; ++split) {
      if (bmode == 1 && mode == 2) continue;
      cudaMemset(dD, 0, D.size() * 4);
      k_test<<<1, 128, smem>>>(dA, dB, dD, mode, split, bmode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d split %d: CUDA error %s\n", mode, split, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
      double e_exact = 0, e_trunc = 0, e_rna = 0, e_lohalf = 0, ref_max = 0;
      for (int m = 0; m < M; ++m)
        for (int n = 0; n < 32; ++n) {
          double ex = 0, tr = 0, rn = 0;
          for (int k = 0; k < K; ++k) {
            const float a = mode == 0 ? A[m * K + k] : A[k * K + m];   // modes 1, 2: D = A^T B
            const float b = B[k * 32 + n];
            ex += (double)a * b;
            tr += (double)trunc_tf32(a) * trunc_tf32(b);
            rn += (double)rna_tf32(a) * rna_tf32(b);
          }
          const double got = split ? (double)D[m * 64 + n] + (double)D[m * 64 + 32 + n] : (double)D[m * 64 + n];
          e_exact = fmax(e_exact, fabs(got - ex)); e_trunc = fmax(e_trunc, fabs(got - tr)); e_rna = fmax(e_rna, fabs(got - rn));
          if (!split) e_lohalf = fmax(e_lohalf, fabs((double)D[m * 64 + 32 + n]));
          ref_max = fmax(ref_max, fabs(ex));
        }
      printf("B %s | mode %d (%s A) split %d: max|D-exact| %.3e  |D-trunc model| %.3e  |D-rna model| %.3e  (ref max %.2f, unsplit lo half %.1e) D[0..2]=%.4f %.4f %.4f\n",
             bmode ? "MN-major" : "K-major", mode, mode == 0 ? "K-major" : (mode == 1 ? "MN-major NONE" : (mode == 2 ? "MN-major SW128_32B" : "A^T in TMEM (TS)")), split, e_exact, e_trunc, e_rna, ref_max, e_lohalf, D[0], D[1], D[2]);
    }
  printf("UMMA_TEST_DONE\n");
  return 0;
}
