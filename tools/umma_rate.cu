// Micro-benchmark (run on the GPU box): cycles per tcgen05.mma kind::tf32 (M = 128, K = 8 per instruction) as a function
// of N and of where the A operand lives (shared memory descriptor vs tensor memory), issued back-to-back by one thread.
// Operand contents are irrelevant (zero-filled shared memory, uninitialised TMEM).
#include <cstdio>
#include <cstdlib>
#include "../mc-gra_b200/csrc/tc_common.cuh"

constexpr uint32_t SJ = 2048 + 16;

__global__ void __launch_bounds__(128) k_rate(long long* out, int n_mma, int N, int ts, int distinct, int sw, int M, int f16, int amn) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ uint32_t tmem_base;
  __shared__ uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < (32 * (int)SJ + 32 * 4112) / 4; e += 128) reinterpret_cast<uint32_t*>(sm)[e] = 0;
  if (warp == 0) tc::tmem_alloc(&tmem_base, 512);
  if (tid == 0) tc::mbar_init(&bar, 1);
  tc::fence_async_smem();
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tm = tmem_base;
  if (tid == 0) {
    const uint32_t lbo_b = (uint32_t)N * 16u + 16u;
    // sw = 1: SWIZZLE_128B K-major operands (rows of 128 B, 8-row atoms of 1024 B, +32 B per K step inside the atom)
    const uint64_t a0 = sw ? tc::make_desc_sw(tc::smem_u32(sm), 16u, 1024u, 2u) : tc::make_desc(tc::smem_u32(sm), SJ, 128u);
    const uint64_t b0 = sw ? tc::make_desc_sw(tc::smem_u32(sm + 32 * 1024), 16u, 1024u, 2u) : tc::make_desc(tc::smem_u32(sm + 32 * SJ), lbo_b, 128u);
    uint32_t idesc = tc::make_idesc_tf32(M, N, 0, 0);
    if (f16) idesc = (idesc & ~((7u << 7) | (7u << 10))) | ((uint32_t)amn << 15);   // a/b format = F16, optional MN-major A
    const uint64_t a_mn = tc::make_desc(tc::smem_u32(sm), 128u, SJ);
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const int ks = distinct ? (i & 3) : 0;
      const uint64_t da = sw ? (uint64_t)(ks * 2) : (uint64_t)((uint32_t)ks * ((2u * SJ) >> 4));
      const uint64_t db = sw ? (uint64_t)(ks * 2) : (uint64_t)((uint32_t)ks * ((2u * lbo_b) >> 4));
      if (f16) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm), "l"(amn ? a_mn + (uint64_t)(ks * 16) : a0 + da), "l"(b0 + db), "r"(idesc), "r"(1u) : "memory");
      } else if (ts) tc::mma_tf32_ts(tm, tm + 256 + ks * 8, b0 + db, idesc, 1u);
      else tc::mma_tf32(tm, a0 + da, b0 + db, idesc, 1u);
    }
    const long long t1 = clock64();
    tc::mma_commit(&bar);
    tc::mbar_wait(&bar, 0);
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tm, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  const size_t smem = 32 * SJ + 32 * 4112 + 1024;
  cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int Ns[] = {32, 64, 128, 256};
  for (int f16 = 0; f16 < 2; ++f16)
  for (int M = 128; M <= 128; M += 64)
  for (int sw = 0; sw < 1 + f16; ++sw)
  for (int ts = 0; ts < 2; ++ts)
    for (int N : Ns) {
        if (f16 && ts) continue;
        const int amn = f16 ? sw : 0;
        const int distinct = 1;
        const int n_mma = 2048;
        k_rate<<<1, 128, smem>>>(d, n_mma, N, ts, distinct, f16 ? 0 : sw, M, f16, amn);
        long long h[2];
        cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        printf("%s M=%3d %s A in %s  N=%3d: issue %.1f clk/MMA, complete %.1f clk/MMA  (floor max(M,128)*N/256 = %d)\n",
               f16 ? "f16  K=16" : "tf32 K=8 ", M, f16 ? (amn ? "A MN-major" : "A K-major ") : (sw ? "SW128" : "NOSWZ"), ts ? "TMEM" : "smem", N, (double)h[0] / n_mma, (double)h[1] / n_mma, N / 2);
      }
  return 0;
}
