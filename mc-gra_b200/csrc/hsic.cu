// HSIC / CKA family on m x d sample matrices without storing m x m kernels.
//
// Reference surface replaced: MC-GRA/hsic.py (distmat :20-27, kernelmat :30-47, hsic_regular
// :117-124, hsic_normalized :127-135, distcorr :50-53, mmd :69-90, mmd_pxpy_pxy :93-114, compute_kernel :56-66),
// utils.HSIC + GaussianKernelMatrix (MC-GRA/utils.py:803-822) and CudaCKA.{rbf, kernel_HSIC, kernel_CKA, linear_HSIC,
// linear_CKA} (MC-GRA/utils.py:1056-1097).  Every Gaussian statistic reduces to five sums over the m^2 pairs,
//     sum K.L,  K1 (row sums),  L1,  1'K1,  1'L1        with  tr(K H L H) = sum K.L - (2/m) K1.L1 + (1'K1)(1'L1)/m^2,
// which one fused pass produces from the Gram tiles (squared distances from 64x64 tiles of X, exp epilogue, tile-wise
// reductions); the linear statistics reduce to weighted cross moments of the factors, O(m d d').
#include "common.cuh"

namespace {

constexpr int GT = 64;          // pair tile edge
constexpr int DMAX = 64;        // max feature width handled by the pair kernels

// rowK[i] += sum_j K_ij, rowL[i] += sum_j L_ij, out[0] += sum_ij K_ij L_ij, out[1] += sum K, out[2] += sum L
__global__ void __launch_bounds__(256)
k_gauss_stats(const float* __restrict__ X, int dx, const float* __restrict__ Y, int dy, int64_t m, float gx, float gy,
              float* __restrict__ rowK, float* __restrict__ rowL, double* __restrict__ out) {
  extern __shared__ float smf[];
  float* xi = smf;                         // [GT][dx+1]
  float* xj = xi + GT * (dx + 1);
  float* yi = xj + GT * (dx + 1);          // [GT][dy+1]
  float* yj = yi + GT * (dy + 1);
  __shared__ double red[32];
  const int64_t i0 = (int64_t)blockIdx.y * GT, j0 = (int64_t)blockIdx.x * GT;
  const int tid = threadIdx.x;
  for (int e = tid; e < GT * dx; e += 256) {
    const int a = e / dx, k = e % dx;
    xi[a * (dx + 1) + k] = (i0 + a < m) ? X[(i0 + a) * dx + k] : 0.f;
    xj[a * (dx + 1) + k] = (j0 + a < m) ? X[(j0 + a) * dx + k] : 0.f;
  }
  for (int e = tid; e < GT * dy; e += 256) {
    const int a = e / dy, k = e % dy;
    yi[a * (dy + 1) + k] = (i0 + a < m) ? Y[(i0 + a) * dy + k] : 0.f;
    yj[a * (dy + 1) + k] = (j0 + a < m) ? Y[(j0 + a) * dy + k] : 0.f;
  }
  __syncthreads();
  // thread (ty, tx): rows ty*4..+3, cols tx*4..+3 of the 64x64 pair tile
  const int tx = tid & 15, ty = tid >> 4;
  float dk[4][4], dl[4][4];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < 4; ++q) dk[p][q] = dl[p][q] = 0.f;
  for (int k = 0; k < dx; ++k) {
    float a[4], b[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) { a[p] = xi[(ty * 4 + p) * (dx + 1) + k]; b[p] = xj[(tx * 4 + p) * (dx + 1) + k]; }
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 4; ++q) { const float df = a[p] - b[q]; dk[p][q] = fmaf(df, df, dk[p][q]); }
  }
  for (int k = 0; k < dy; ++k) {
    float a[4], b[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) { a[p] = yi[(ty * 4 + p) * (dy + 1) + k]; b[p] = yj[(tx * 4 + p) * (dy + 1) + k]; }
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 4; ++q) { const float df = a[p] - b[q]; dl[p][q] = fmaf(df, df, dl[p][q]); }
  }
  double skl = 0.0, sk = 0.0, sl = 0.0;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int64_t gi = i0 + ty * 4 + p;
    float rk = 0.f, rl = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int64_t gj = j0 + tx * 4 + q;
      if (gi < m && gj < m) {
        const float kv = expf(-dk[p][q] * gx), lv = expf(-dl[p][q] * gy);
        rk += kv; rl += lv;
        skl += (double)(kv * lv);
      }
    }
    sk += rk; sl += rl;
    // reduce the row partials over the 16 tx lanes of this ty (a half-warp)
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      rk += __shfl_xor_sync(0xffffffffu, rk, o);
      rl += __shfl_xor_sync(0xffffffffu, rl, o);
    }
    if (tx == 0 && gi < m) { atomicAdd(rowK + gi, rk); atomicAdd(rowL + gi, rl); }
  }
  block_atomic_add_d(skl, out + 0, red);
  block_atomic_add_d(sk, out + 1, red);
  block_atomic_add_d(sl, out + 2, red);
}

// dense pair matrix between two point sets: mode 0 squared distance, 1 exp(-gamma * sqdist)
__global__ void __launch_bounds__(256)
k_pair_dense(const float* __restrict__ X, int d, int64_t m1, const float* __restrict__ Z, int64_t m2, int mode, float gamma,
             float* __restrict__ out) {
  extern __shared__ float smf[];
  float* xi = smf;
  float* zj = xi + 32 * (d + 1);
  const int64_t i0 = (int64_t)blockIdx.y * 32, j0 = (int64_t)blockIdx.x * 32;
  const int tid = threadIdx.x;
  for (int e = tid; e < 32 * d; e += 256) {
    const int a = e / d, k = e % d;
    xi[a * (d + 1) + k] = (i0 + a < m1) ? X[(i0 + a) * d + k] : 0.f;
    zj[a * (d + 1) + k] = (j0 + a < m2) ? Z[(j0 + a) * d + k] : 0.f;
  }
  __syncthreads();
  const int tx = tid & 31, ty = tid >> 5;
  for (int rr = ty; rr < 32; rr += 8) {
    const int64_t gi = i0 + rr, gj = j0 + tx;
    if (gi >= m1 || gj >= m2) continue;
    float s = 0.f;
    for (int k = 0; k < d; ++k) { const float df = xi[rr * (d + 1) + k] - zj[tx * (d + 1) + k]; s = fmaf(df, df, s); }
    out[gi * m2 + gj] = mode == 0 ? s : expf(-gamma * s);
  }
}

// weighted raw moments: out = [ sum w x (dx) | sum w y (dy) | sum w x y^T (dx*dy) | sum w y y^T (dy*dy) ] in fp64
__global__ void __launch_bounds__(256)
k_cross_moments(const float* __restrict__ X, int dx, const float* __restrict__ Y, int dy, const float* __restrict__ w,
                int64_t n, double* __restrict__ out) {
  extern __shared__ float smf[];
  float* xs = smf;                 // [128][dx]
  float* ys = xs + 128 * dx;       // [128][dy]
  float* ws = ys + 128 * dy;       // [128]
  const int64_t i0 = (int64_t)blockIdx.x * 128;
  const int tid = threadIdx.x;
  for (int e = tid; e < 128 * dx; e += 256) xs[e] = (i0 + e / dx < n) ? X[(i0 + e / dx) * dx + e % dx] : 0.f;
  for (int e = tid; e < 128 * dy; e += 256) ys[e] = (i0 + e / dy < n) ? Y[(i0 + e / dy) * dy + e % dy] : 0.f;
  if (tid < 128) ws[tid] = (i0 + tid < n) ? (w ? w[i0 + tid] : 1.f) : 0.f;
  __syncthreads();
  const int total = dx + dy + dx * dy + dy * dy;
  for (int q = tid; q < total; q += 256) {
    double s = 0.0;
    if (q < dx) {
      for (int a = 0; a < 128; ++a) s += (double)(ws[a] * xs[a * dx + q]);
    } else if (q < dx + dy) {
      const int k = q - dx;
      for (int a = 0; a < 128; ++a) s += (double)(ws[a] * ys[a * dy + k]);
    } else if (q < dx + dy + dx * dy) {
      const int k = (q - dx - dy) / dy, l = (q - dx - dy) % dy;
      for (int a = 0; a < 128; ++a) s += (double)(ws[a] * xs[a * dx + k]) * (double)ys[a * dy + l];
    } else {
      const int k = (q - dx - dy - dx * dy) / dy, l = (q - dx - dy - dx * dy) % dy;
      for (int a = 0; a < 128; ++a) s += (double)(ws[a] * ys[a * dy + k]) * (double)ys[a * dy + l];
    }
    if (s != 0.0) atomicAdd(out + q, s);
  }
}

// backward of k_cross_moments: g = d loss / d out (same layout, fp64);
//   dX[i] = w_i (g_sx + g_Sxy y_i),   dY[i] = w_i (g_sy + g_Sxy^T x_i + (g_Syy + g_Syy^T) y_i)
__global__ void k_cross_moments_bwd(const float* __restrict__ X, int dx, const float* __restrict__ Y, int dy,
                                    const float* __restrict__ w, int64_t n, const double* __restrict__ g,
                                    float* __restrict__ dX, float* __restrict__ dY) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double wi = w ? (double)w[i] : 1.0;
  const double* gsx = g;
  const double* gsy = g + dx;
  const double* gxy = g + dx + dy;
  const double* gyy = gxy + dx * dy;
  if (dX != nullptr) {
    for (int k = 0; k < dx; ++k) {
      double s = gsx[k];
      for (int l = 0; l < dy; ++l) s += gxy[k * dy + l] * (double)Y[i * dy + l];
      dX[i * dx + k] = (float)(wi * s);
    }
  }
  if (dY != nullptr) {
    for (int l = 0; l < dy; ++l) {
      double s = gsy[l];
      for (int k = 0; k < dx; ++k) s += gxy[k * dy + l] * (double)X[i * dx + k];
      for (int m = 0; m < dy; ++m) s += (gyy[l * dy + m] + gyy[m * dy + l]) * (double)Y[i * dy + m];
      dY[i * dy + l] = (float)(wi * s);
    }
  }
}

}  // namespace

extern "C" {

int mcgra_cross_moments_bwd(const float* X, int dx, const float* Y, int dy, const float* w, int64_t n, const double* g,
                            float* dX, float* dY, void* stream) {
  if (dx < 1 || dy < 1 || dx > 64 || dy > 64) return -1;
  if (n <= 0) return 0;
  k_cross_moments_bwd<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(X, dx, Y, dy, w, n, g, dX, dY);
  MCGRA_LAUNCH_CHECK();
  return 0;
}

int mcgra_gauss_stats(const float* X, int dx, const float* Y, int dy, int64_t m, float gx, float gy, float* rowK,
                      float* rowL, double* out, void* stream) {
  if (dx < 1 || dy < 1 || dx > DMAX || dy > DMAX) return -1;
  if (m <= 0) return 0;
  dim3 grid((unsigned)((m + GT - 1) / GT), (unsigned)((m + GT - 1) / GT));
  if (grid.y > 65535) return -3;
  const size_t smem = (size_t)(2 * GT * (dx + 1) + 2 * GT * (dy + 1)) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(k_gauss_stats, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  k_gauss_stats<<<grid, 256, smem, (cudaStream_t)stream>>>(X, dx, Y, dy, m, gx, gy, rowK, rowL, out);
  MCGRA_LAUNCH_CHECK();
  return 0;
}

int mcgra_pair_dense(const float* X, int d, int64_t m1, const float* Z, int64_t m2, int mode, float gamma, float* out,
                     void* stream) {
  if (d < 1 || d > 4096) return -1;
  if (m1 <= 0 || m2 <= 0) return 0;
  dim3 grid((unsigned)((m2 + 31) / 32), (unsigned)((m1 + 31) / 32));
  if (grid.y > 65535) return -3;
  const size_t smem = (size_t)(2 * 32 * (d + 1)) * sizeof(float);
  if (smem > 200 * 1024) return -1;
  cudaError_t e = cudaFuncSetAttribute(k_pair_dense, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  k_pair_dense<<<grid, 256, smem, (cudaStream_t)stream>>>(X, d, m1, Z, m2, mode, gamma, out);
  MCGRA_LAUNCH_CHECK();
  return 0;
}

int mcgra_cross_moments(const float* X, int dx, const float* Y, int dy, const float* w, int64_t n, double* out,
                        void* stream) {
  if (dx < 1 || dy < 1 || dx > 64 || dy > 64) return -1;
  if (n <= 0) return 0;
  const size_t smem = (size_t)(128 * dx + 128 * dy + 128) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(k_cross_moments, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  k_cross_moments<<<(unsigned)((n + 127) / 128), 256, smem, (cudaStream_t)stream>>>(X, dx, Y, dy, w, n, out);
  MCGRA_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
