// Ceiling probe for the fold kernel's HBM pattern: 4 read streams (x, m, v, F) + 3 write streams (x, m, v), walked in
// 64 KB tile blocks by persistent CTAs, with trivial arithmetic.  Variants:
//   ldg  : per-warp float4 loads/stores (rows of 512 B), W warps per CTA, U rows in flight per warp, B CTAs per SM
//   bulk : cp.async.bulk global->shared ring (S stages of R rows x 4 arrays), compute in place in shared memory,
//          cp.async.bulk shared->global stores
// usage: stream7 [tiles]      (prints GB/s of algorithmic traffic = 7 x 64 KB per tile)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int TILE = 128;
constexpr int TE = TILE * TILE;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int U, bool CS>
__global__ void k_ldg(float* __restrict__ x, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ f, int ntiles) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int rpw = TILE / nw;     // rows per warp per tile
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t base = (int64_t)t * TE + lane * 4;
    for (int it = 0; it < rpw / U; ++it) {
      float4 X[U], M[U], V[U], F[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t o = base + (int64_t)((it * U + u) * nw + warp) * TILE;
        if (CS) {
          X[u] = __ldcs((const float4*)(x + o)); M[u] = __ldcs((const float4*)(m + o));
          V[u] = __ldcs((const float4*)(v + o)); F[u] = __ldcs((const float4*)(f + o));
        } else {
          X[u] = *(const float4*)(x + o); M[u] = *(const float4*)(m + o);
          V[u] = *(const float4*)(v + o); F[u] = *(const float4*)(f + o);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t o = base + (int64_t)((it * U + u) * nw + warp) * TILE;
        float4 a = make_float4(X[u].x + F[u].x, X[u].y + F[u].y, X[u].z + F[u].z, X[u].w + F[u].w);
        float4 b = make_float4(M[u].x * 0.9f + a.x, M[u].y * 0.9f + a.y, M[u].z * 0.9f + a.z, M[u].w * 0.9f + a.w);
        float4 c = make_float4(V[u].x * 0.99f + a.x * a.x, V[u].y * 0.99f + a.y * a.y, V[u].z * 0.99f + a.z * a.z, V[u].w * 0.99f + a.w * a.w);
        if (CS) { __stcs((float4*)(x + o), a); __stcs((float4*)(m + o), b); __stcs((float4*)(v + o), c); }
        else { *(float4*)(x + o) = a; *(float4*)(m + o) = b; *(float4*)(v + o) = c; }
      }
    }
  }
}

// ---- bulk-async variant -------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n\t.reg .pred P1;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra.uni WD;\n\tbra.uni WL;\n\tWD:\n\t}" ::"r"(s32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(s32(src)), "r"(bytes) : "memory");
}

// S stages; each stage holds R rows of x, m, v, f (R * 512 B each).  Thread 0 = producer + store issuer.
template <int S, int R>
__global__ void k_bulk(float* __restrict__ x, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ f, int ntiles) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int CH = R * TILE;                 // floats per array per stage
  float* buf = reinterpret_cast<float*>(smem); // [S][4][CH]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)S * 4 * CH * 4);
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int chunks_per_tile = TILE / R;
  const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total = my_tiles * chunks_per_tile;
  auto issue = [&](int c) {
    const int s = c % S;
    const int t = blockIdx.x + (c / chunks_per_tile) * gridDim.x;
    const int64_t o = (int64_t)t * TE + (int64_t)(c % chunks_per_tile) * CH;
    float* b = buf + (size_t)s * 4 * CH;
    mbar_expect(&full[s], 4u * CH * 4u);
    bulk_g2s(b, x + o, CH * 4, &full[s]);
    bulk_g2s(b + CH, m + o, CH * 4, &full[s]);
    bulk_g2s(b + 2 * CH, v + o, CH * 4, &full[s]);
    bulk_g2s(b + 3 * CH, f + o, CH * 4, &full[s]);
  };
  if (tid == 0)
    for (int c = 0; c < S - 1 && c < total; ++c) issue(c);
  for (int c = 0; c < total; ++c) {
    const int s = c % S;
    // refill the stage that chunk c - 1 used (its stores must have finished READING shared memory)
    if (tid == 0 && c + S - 1 < total) {
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      issue(c + S - 1);
    }
    mbar_wait(&full[s], (uint32_t)((c / S) & 1));
    float* b = buf + (size_t)s * 4 * CH;
    for (int e = tid * 4; e < CH; e += nt * 4) {
      float4 X = *(float4*)(b + e), M = *(float4*)(b + CH + e), V = *(float4*)(b + 2 * CH + e), F = *(float4*)(b + 3 * CH + e);
      float4 a = make_float4(X.x + F.x, X.y + F.y, X.z + F.z, X.w + F.w);
      *(float4*)(b + e) = a;
      *(float4*)(b + CH + e) = make_float4(M.x * 0.9f + a.x, M.y * 0.9f + a.y, M.z * 0.9f + a.z, M.w * 0.9f + a.w);
      *(float4*)(b + 2 * CH + e) = make_float4(V.x * 0.99f + a.x * a.x, V.y * 0.99f + a.y * a.y, V.z * 0.99f + a.z * a.z, V.w * 0.99f + a.w * a.w);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      const int t = blockIdx.x + (c / chunks_per_tile) * gridDim.x;
      const int64_t o = (int64_t)t * TE + (int64_t)(c % chunks_per_tile) * CH;
      bulk_s2g(x + o, b, CH * 4);
      bulk_s2g(m + o, b + CH, CH * 4);
      bulk_s2g(v + o, b + 2 * CH, CH * 4);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename F>
float timeit(F&& launch, int reps = 5) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(a));
    launch();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    best = ms < best ? ms : best;
  }
  CK(cudaGetLastError());
  return best;
}

int main(int argc, char** argv) {
  const int ntiles = argc > 1 ? atoi(argv[1]) : 65536;       // 65536 tiles x 64 KB = 4.3 GB per array
  const size_t bytes = (size_t)ntiles * TE * 4;
  float *x, *m, *v, *f;
  CK(cudaMalloc(&x, bytes)); CK(cudaMalloc(&m, bytes)); CK(cudaMalloc(&v, bytes)); CK(cudaMalloc(&f, bytes));
  CK(cudaMemset(x, 0, bytes)); CK(cudaMemset(m, 0, bytes)); CK(cudaMemset(v, 0, bytes)); CK(cudaMemset(f, 0, bytes));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const double gb = 7.0 * bytes / 1e9;
  printf("tiles %d, %.1f GB per array, algorithmic traffic %.1f GB, %d SMs\n", ntiles, bytes / 1e9, gb, sms);
  {  // plain copy reference: x -> m (read 1, write 1)
    float ms = timeit([&] { CK(cudaMemcpyAsync(m, x, bytes, cudaMemcpyDeviceToDevice)); });
    printf("cudaMemcpy d2d            : %.3f ms  %.0f GB/s (read+write)\n", ms, 2.0 * bytes / 1e6 / ms);
  }
#define RUN_LDG(U, CS, W, B)                                                                          \
  {                                                                                                   \
    float ms = timeit([&] { k_ldg<U, CS><<<sms * B, W * 32>>>(x, m, v, f, ntiles); });                \
    printf("ldg U=%d cs=%d warps=%2d cta/sm=%d : %.3f ms  %.0f GB/s\n", U, (int)CS, W, B, ms, gb * 1e3 / ms); \
  }
  RUN_LDG(4, false, 8, 2)
  RUN_LDG(4, true, 8, 2)
  RUN_LDG(4, false, 8, 1)
  RUN_LDG(4, false, 16, 1)
  RUN_LDG(2, false, 16, 2)
  RUN_LDG(8, false, 8, 1)
  RUN_LDG(8, false, 8, 2)
  RUN_LDG(4, false, 8, 4)
  RUN_LDG(2, false, 8, 8)
  RUN_LDG(4, true, 8, 4)
#define RUN_BULK(S, R, T, B)                                                                          \
  {                                                                                                   \
    const size_t sm = (size_t)S * 4 * R * TILE * 4 + 64;                                              \
    CK(cudaFuncSetAttribute(k_bulk<S, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));     \
    float ms = timeit([&] { k_bulk<S, R><<<sms * B, T, sm>>>(x, m, v, f, ntiles); });                 \
    printf("bulk S=%d R=%2d threads=%d cta/sm=%d (%zu KB smem): %.3f ms  %.0f GB/s\n", S, R, T, B, sm / 1024, ms, gb * 1e3 / ms); \
  }
  RUN_BULK(3, 16, 256, 1)
  RUN_BULK(4, 16, 256, 1)
  RUN_BULK(3, 16, 256, 2)
  RUN_BULK(2, 32, 256, 1)
  RUN_BULK(3, 32, 512, 1)
  RUN_BULK(4, 8, 256, 2)
  RUN_BULK(6, 8, 256, 1)
  RUN_BULK(3, 8, 128, 4)
  return 0;
}
