// O(n^2) helpers around the dense contraction mcgra_gemm_nt (gemm.cu): building the fp16x2 operand images of
// A_hat (utils.normalize_adj_tensor, utils.py:211-230), M1 (get_modified_adj_after of dot_product_decode,
// topology_attack.py:381-395, 414-419), of dense fp32 matrices and their transposes; CudaCKA.centering
// (utils.py:1060-1065) in closed form; the rank-1 vectors of the centred products; the scalar algebra of
// linear_HSIC / linear_CKA / dot_product (utils.py:1080-1091, topology_attack.py:480-481) and the hand-over of the
// dense gradients to the tiled triangle.  All streaming, HBM-bound.
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

// power-of-two row scale from the bits of a non-negative bound m >= max_j |x_ij|:  m * s <= 2^14
__device__ __forceinline__ void scale_from_bits(uint32_t bits, float& s, float& inv_s) {
  const int eb = (int)((bits >> 23) & 0xffu);             // m < 2^(eb - 126)
  int se = (bits == 0u) ? 0 : (140 - eb);                 // s = 2^se
  se = max(-100, min(100, se));
  s = __int_as_float((127 + se) << 23);
  inv_s = __int_as_float((127 - se) << 23);
}
__device__ __forceinline__ void split_store(float xs, __half* hi, __half* lo) {
  const __half h = __float2half_rn(xs);
  *hi = h;
  *lo = __float2half_rn(xs - __half2float(h));
}

// ---- image of a dense matrix -----------------------------------------------------------------------------------
__global__ void k_rowabsmax(const float* __restrict__ X, int64_t rows, int64_t cols, int64_t ld, uint32_t* __restrict__ mx) {
  const int64_t i = blockIdx.x;
  float m = 0.f;
  for (int64_t j = threadIdx.x; j < cols; j += blockDim.x) m = fmaxf(m, fabsf(X[i * ld + j]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(mx + i, __float_as_uint(m));
}
__global__ void k_colabsmax(const float* __restrict__ X, int64_t rows, int64_t cols, int64_t ld, uint32_t* __restrict__ mx) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.y * 256, r1 = min(rows, r0 + 256);
  if (j >= cols) return;
  float m = 0.f;
  for (int64_t i = r0; i < r1; ++i) m = fmaxf(m, fabsf(X[i * ld + j]));
  if (m > 0.f) atomicMax(mx + j, __float_as_uint(m));
}
__global__ void k_image_rows(const float* __restrict__ X, int64_t rows, int64_t cols, int64_t ld,
                             const uint32_t* __restrict__ mx, __half* __restrict__ hi, __half* __restrict__ lo,
                             float* __restrict__ inv_scale, int64_t ldo) {
  const int64_t i = blockIdx.x;
  float s, is;
  scale_from_bits(mx[i], s, is);
  if (threadIdx.x == 0) inv_scale[i] = is;
  for (int64_t j = threadIdx.x; j < cols; j += blockDim.x) split_store(X[i * ld + j] * s, hi + i * ldo + j, lo + i * ldo + j);
}
// out row o = column o of X (32 x 32 shared-memory transpose)
__global__ void k_image_cols(const float* __restrict__ X, int64_t rows, int64_t cols, int64_t ld,
                             const uint32_t* __restrict__ mx, __half* __restrict__ hi, __half* __restrict__ lo,
                             float* __restrict__ inv_scale, int64_t ldo) {
  __shared__ float t[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;        // 32 x 8
  const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  for (int rr = ty; rr < 32; rr += 8) {
    const int64_t i = r0 + rr, j = c0 + tx;
    t[rr][tx] = (i < rows && j < cols) ? X[i * ld + j] : 0.f;
  }
  __syncthreads();
  for (int rr = ty; rr < 32; rr += 8) {
    const int64_t o = c0 + rr, k = r0 + tx;                      // output row o (column of X), output column k (row of X)
    if (o < cols && k < rows) {
      float s, is;
      scale_from_bits(mx[o], s, is);
      split_store(t[tx][rr] * s, hi + o * ldo + k, lo + o * ldo + k);
      if (k == 0) inv_scale[o] = is;
    }
  }
}

// ---- image of A_hat from the full tiled triangle -----------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_image_ahat(const float* __restrict__ tiles, int64_t n, const float* mu, int raw, const float* __restrict__ r,
             __half* __restrict__ hi, __half* __restrict__ lo, float* __restrict__ inv_scale, int64_t ldo,
             double* __restrict__ rowsum) {
  __shared__ float t[32][33];
  const int I = blockIdx.y, J = blockIdx.x;
  const ParamView pv = load_view(mu, raw);
  const bool need_d = J <= I, need_m = J >= I;
  const float* src_d = need_d ? tiles + (tri((int64_t)I) + J) * TILE_ELEMS : nullptr;
  const float* src_m = need_m ? tiles + (tri((int64_t)J) + I) * TILE_ELEMS : nullptr;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int br = 0; br < TILE; br += 32) {
    for (int bc = 0; bc < TILE; bc += 32) {
      __syncthreads();
      if (need_m)
        for (int rr = ty; rr < 32; rr += 8) t[rr][tx] = src_m[(bc + rr) * TILE + br + tx];
      __syncthreads();
      for (int rr = ty; rr < 32; rr += 8) {
        const int a = br + rr, b = bc + tx;
        const int64_t i = (int64_t)I * TILE + a, j = (int64_t)J * TILE + b;
        float val = 0.f;
        if (i < n && j < n) {
          const float ri = r[i];
          if (i == j) {
            val = ri * ri;
          } else {
            const float xs = (j < i) ? src_d[a * TILE + b] : t[tx][rr];
            val = (ri * pv.adj(xs)) * r[j];                      // value order of utils.py:227-229
          }
          float s, is;
          scale_from_bits(__float_as_uint(ri), s, is);
          split_store(val * s, hi + i * ldo + j, lo + i * ldo + j);
          if (j == 0) inv_scale[i] = is;
        }
        const float rs = warp_sum(val);
        if (tx == 0 && i < n && rs != 0.f) atomicAdd(rowsum + i, (double)rs);
      }
    }
  }
}

// ---- image of M1 = relu(zhat zhat^T), zero diagonal --------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_image_m1(const float* __restrict__ zhat, int64_t n, __half* __restrict__ hi, __half* __restrict__ lo,
           float* __restrict__ inv_scale, int64_t ldo, double* __restrict__ rowsum) {
  __shared__ float zi[TILE][HID + 1], zj[TILE][HID + 1];
  const int I = blockIdx.y, J = blockIdx.x;
  for (int e = threadIdx.x; e < TILE * HID; e += blockDim.x) {
    const int a = e / HID, k = e % HID;
    const int64_t i = (int64_t)I * TILE + a, j = (int64_t)J * TILE + a;
    zi[a][k] = i < n ? zhat[i * HID + k] : 0.f;
    zj[a][k] = j < n ? zhat[j * HID + k] : 0.f;
  }
  __syncthreads();
  const int cp = threadIdx.x & 63, rs = threadIdx.x >> 6;       // column pair, row sub-index
  const float S = 16384.f;
  for (int a = rs; a < TILE; a += 4) {
    const int64_t i = (int64_t)I * TILE + a;
    float v[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int b = 2 * cp + h;
      const int64_t j = (int64_t)J * TILE + b;
      float d = 0.f;
#pragma unroll
      for (int k = 0; k < HID; ++k) d = fmaf(zi[a][k], zj[b][k], d);
      v[h] = (i < n && j < n && i != j) ? fmaxf(d, 0.f) : 0.f;
      if (i < n && j < n) split_store(v[h] * S, hi + i * ldo + j, lo + i * ldo + j);
    }
    const float sum = warp_sum(v[0] + v[1]);
    if ((threadIdx.x & 31) == 0 && i < n && sum != 0.f) atomicAdd(rowsum + i, (double)sum);
    if (J == 0 && cp == 0 && i < n) inv_scale[i] = 1.f / S;
  }
}

// ---- centring H X H of a symmetric matrix -----------------------------------------------------------------------
__global__ void k_rowsum_d(const float* __restrict__ X, int64_t n, int64_t ld, double* __restrict__ ws) {
  __shared__ double red[32];
  const int64_t i = blockIdx.x;
  double s = 0.0;
  for (int64_t j = threadIdx.x; j < n; j += blockDim.x) s += (double)X[i * ld + j];
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    v = warp_sum_d(v);
    if (threadIdx.x == 0) {
      ws[i] = v;
      atomicAdd(ws + n, v);
    }
  }
}
__global__ void k_center_apply(float* __restrict__ X, int64_t n, int64_t ld, const double* __restrict__ ws) {
  const int64_t i = blockIdx.x;
  const double inv = 1.0 / (double)n;
  const double mi = ws[i] * inv, tot = ws[n] * inv * inv;
  for (int64_t j = threadIdx.x; j < n; j += blockDim.x)
    X[i * ld + j] = (float)((double)X[i * ld + j] - mi - ws[j] * inv + tot);
}

__global__ void k_sumsq_d(const float* __restrict__ X, int64_t cols, int64_t ld, double* __restrict__ out) {
  __shared__ double red[32];
  const int64_t i = blockIdx.x;
  double s = 0.0;
  for (int64_t j = threadIdx.x; j < cols; j += blockDim.x) {
    const double v = (double)X[i * ld + j];
    s += v * v;
  }
  block_atomic_add_d(s, out, red);
}

// ---- GEMV helpers -------------------------------------------------------------------------------------------------
__global__ void k_gemv_cols(const float* __restrict__ X, int64_t rows, int64_t cols, int64_t ld, const double* __restrict__ w,
                            double scale, double* __restrict__ out) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.y * 256, r1 = min(rows, r0 + 256);
  if (j >= cols) return;
  double s = 0.0;
  for (int64_t k = r0; k < r1; ++k) s += w[k] * (double)X[k * ld + j];
  atomicAdd(out + j, s * scale);
}
__global__ void k_gemv_rows(const float* __restrict__ X, int64_t rows, int64_t cols, int64_t ld, const double* __restrict__ w,
                            double scale, double* __restrict__ out) {
  const int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= rows) return;
  double s = 0.0;
  for (int64_t k = threadIdx.x & 31; k < cols; k += 32) s += (double)X[j * ld + k] * w[k];
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) out[j] += s * scale;
}

// ---- dense gradient -> tiled triangle ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_sym_to_tiles(const float* __restrict__ G, int64_t ld, int64_t n, int64_t t0, float scale, const float* scale_dev,
               float* __restrict__ tiles) {
  __shared__ float t[32][33];
  int I, J;
  tile_coords(t0 + blockIdx.x, I, J);
  const float sc = scale * (scale_dev != nullptr ? *scale_dev : 1.f);
  float* dst = tiles + (int64_t)blockIdx.x * TILE_ELEMS;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int br = 0; br < TILE; br += 32) {
    for (int bc = 0; bc < TILE; bc += 32) {
      __syncthreads();
      for (int rr = ty; rr < 32; rr += 8) {                     // block of G^T: rows J*128+bc.., cols I*128+br..
        const int64_t gi = (int64_t)J * TILE + bc + rr, gj = (int64_t)I * TILE + br + tx;
        t[rr][tx] = (gi < n && gj < n) ? G[gi * ld + gj] : 0.f;
      }
      __syncthreads();
      for (int rr = ty; rr < 32; rr += 8) {
        const int a = br + rr, b = bc + tx;
        const int64_t i = (int64_t)I * TILE + a, j = (int64_t)J * TILE + b;
        float v = 0.f;
        if (j < i && i < n) v = (G[i * ld + j] + t[tx][rr]) * sc;
        dst[a * TILE + b] = v;
      }
    }
  }
}
__global__ void k_diag_scaled(const float* __restrict__ G, int64_t ld, int64_t n, float scale, const float* scale_dev,
                              float* __restrict__ diag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) diag[i] = G[i * ld + i] * scale * (scale_dev != nullptr ? *scale_dev : 1.f);
}

// ---- scalar algebra of the measures ----------------------------------------------------------------------------------
__global__ void k_dense_scalars(int measure, const double* __restrict__ in, double k1c, double k2c, double sign,
                                double* __restrict__ acc, float* __restrict__ alpha) {
  const double S1 = in[0], hAM = in[1], hAA = in[2], hMM = in[3], hFF = in[4];
  double c1 = 0.0, c2 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, a4 = 0.0, a5 = 0.0;
  if (measure == MCGRA_M_HSIC) {
    c1 = k1c * S1;
    c2 = k2c * hAM;
    a1 = 2.0 * k1c;
    a2 = a4 = 2.0 * k2c;
  } else if (measure == MCGRA_M_DP) {              // || Y^T X ||_F = sqrt(.)
    const double n1 = sqrt(fmax(S1, 0.0)), n2 = sqrt(fmax(hAM, 0.0));
    c1 = k1c * n1;
    c2 = k2c * n2;
    a1 = n1 > 0.0 ? k1c / n1 : 0.0;
    a2 = a4 = n2 > 0.0 ? k2c / n2 : 0.0;
  } else if (measure == MCGRA_M_CKA) {             // hsic / (sqrt(hsic_xx) sqrt(hsic_yy))
    if (k1c != 0.0) {
      const double den = sqrt(hFF) * sqrt(hAA);
      c1 = k1c * S1 / den;
      a1 = 2.0 * k1c / den;
      a3 += -2.0 * c1 / hAA;
    }
    if (k2c != 0.0) {
      const double den = sqrt(hAA) * sqrt(hMM);
      c2 = k2c * hAM / den;
      a2 = a4 = 2.0 * k2c / den;
      a3 += -2.0 * c2 / hAA;
      a5 = -2.0 * c2 / hMM;
    }
  }
  acc[MCGRA_ACC_C1D] += c1;
  acc[MCGRA_ACC_C2D] += c2;
  alpha[0] = (float)(sign * a1);
  alpha[1] = (float)(sign * a2);
  alpha[2] = (float)(sign * a3);
  alpha[3] = (float)(sign * a4);
  alpha[4] = (float)(sign * a5);
  alpha[5] = 1.f;
}

__global__ void k_d2f(const double* __restrict__ in, int64_t count, double scale, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = (float)(in[i] * scale);
}

}  // namespace

extern "C" {

int mcgra_image_from_dense(const float* src, int64_t rows, int64_t cols, int64_t ld, int transpose,
                           const mcgra_image* out, void* ws, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (src == nullptr || out == nullptr || ws == nullptr) return -1;
  const int64_t orows = transpose ? cols : rows, ocols = transpose ? rows : cols;
  if (out->rows != orows || out->cols != ocols || out->ld < ocols) return -2;
  if (rows <= 0 || cols <= 0) return 0;
  uint32_t* mx = (uint32_t*)ws;
  cudaError_t e = cudaMemsetAsync(mx, 0, sizeof(uint32_t) * (size_t)orows, st);
  if (e != cudaSuccess) return (int)e;
  if (!transpose) {
    k_rowabsmax<<<(unsigned)rows, 256, 0, st>>>(src, rows, cols, ld, mx);
    k_image_rows<<<(unsigned)rows, 256, 0, st>>>(src, rows, cols, ld, mx, (__half*)out->hi, (__half*)out->lo,
                                                 out->inv_scale, out->ld);
  } else {
    dim3 g1((unsigned)((cols + 255) / 256), (unsigned)((rows + 255) / 256));
    k_colabsmax<<<g1, 256, 0, st>>>(src, rows, cols, ld, mx);
    dim3 g2((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
    k_image_cols<<<g2, 256, 0, st>>>(src, rows, cols, ld, mx, (__half*)out->hi, (__half*)out->lo, out->inv_scale, out->ld);
  }
  MCGRA_LAUNCH_CHECK();
  return 0;
}

int mcgra_image_ahat(const float* tiles, int64_t n, const float* mu, int raw, const float* r, const mcgra_image* out,
                     double* rowsum, void* stream) {
  if (out == nullptr || out->rows != n || out->cols != n) return -2;
  const unsigned T = (unsigned)((n + TILE - 1) / TILE);
  k_image_ahat<<<dim3(T, T), 256, 0, (cudaStream_t)stream>>>(tiles, n, mu, raw, r, (__half*)out->hi, (__half*)out->lo,
                                                              out->inv_scale, out->ld, rowsum);
  MCGRA_LAUNCH_CHECK();
  return 0;
}

int mcgra_image_m1(const float* zhat, int64_t n, const mcgra_image* out, double* rowsum, void* stream) {
  if (out == nullptr || out->rows != n || out->cols != n) return -2;
  const unsigned T = (unsigned)((n + TILE - 1) / TILE);
  k_image_m1<<<dim3(T, T), 256, 0, (cudaStream_t)stream>>>(zhat, n, (__half*)out->hi, (__half*)out->lo, out->inv_scale,
                                                            out->ld, rowsum);
  MCGRA_LAUNCH_CHECK();
  return 0;
}

int mcgra_center_dense(float* X, int64_t n, int64_t ld, double* ws, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(ws + n, 0, sizeof(double), st);
  if (e != cudaSuccess) return (int)e;
  k_rowsum_d<<<(unsigned)n, 256, 0, st>>>(X, n, ld, ws);
  k_center_apply<<<(unsigned)n, 256, 0, st>>>(X, n, ld, ws);
  MCGRA_LAUNCH_CHECK();
  return 0;
}

int mcgra_dense_gemv(const float* X, int64_t rows, int64_t cols, int64_t ld, const double* w, double scale,
                     int transpose, double* out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!transpose) {
    dim3 g((unsigned)((cols + 255) / 256), (unsigned)((rows + 255) / 256));
    k_gemv_cols<<<g, 256, 0, st>>>(X, rows, cols, ld, w, scale, out);
  } else {
    k_gemv_rows<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(X, rows, cols, ld, w, scale, out);
  }
  MCGRA_LAUNCH_CHECK();
  return 0;
}

int mcgra_dense_sumsq(const float* X, int64_t rows, int64_t cols, int64_t ld, double* out, void* stream) {
  if (rows <= 0) return 0;
  k_sumsq_d<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(X, cols, ld, out);
  MCGRA_LAUNCH_CHECK();
  return 0;
}

int mcgra_sym_to_tiles(const float* G, int64_t ld, int64_t n, int tr0, int tr1, float scale, const float* scale_dev,
                       float* tiles, float* diag, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nt = tri(tr1) - tri(tr0);
  if (nt > 0) k_sym_to_tiles<<<(unsigned)nt, 256, 0, st>>>(G, ld, n, tri(tr0), scale, scale_dev, tiles);
  if (diag != nullptr) k_diag_scaled<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(G, ld, n, scale, scale_dev, diag);
  MCGRA_LAUNCH_CHECK();
  return 0;
}

int mcgra_dense_scalars(int measure, const double* in, double k1c, double k2c, double sign, double* acc, float* alpha,
                        void* stream) {
  k_dense_scalars<<<1, 1, 0, (cudaStream_t)stream>>>(measure, in, k1c, k2c, sign, acc, alpha);
  MCGRA_LAUNCH_CHECK();
  return 0;
}

int mcgra_d2f(const double* in, int64_t count, double scale, float* out, void* stream) {
  if (count <= 0) return 0;
  k_d2f<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, count, scale, out);
  MCGRA_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
