// Stand-alone helpers of the reference class surface (topology_attack.py:469-487): adding_noise's add + clamp and
// calc_kl's row-softmax KL divergence on arbitrary 2-D operands.  Streaming, one pass each.
#include "common.cuh"

namespace {

// M <- clamp(M + eps * noise, 0, 1)   (adding_noise, :474-478; the N(0,1) draw itself is the caller's RNG stream)
__global__ void k_noise_clamp(float* __restrict__ M, const float* __restrict__ noise, float eps, int64_t count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) M[i] = fminf(fmaxf(fmaf(noise[i], eps, M[i]), 0.f), 1.f);
}

// out += sum_i sum_j p_ij (log p_ij - log q_ij),  p = softmax(X_i), q = softmax(Y_i)   (calc_kl, :483-487; the caller
// divides by the row count for reduction="batchmean").  One CTA per row, max-shifted like F.softmax / F.log_softmax.
__global__ void __launch_bounds__(256)
k_row_kl(const float* __restrict__ X, const float* __restrict__ Y, int64_t cols, int64_t ldx, int64_t ldy,
         double* __restrict__ out) {
  __shared__ float redf[32];
  __shared__ double redd[32];
  const int64_t i = blockIdx.x;
  const float* x = X + i * ldx;
  const float* y = Y + i * ldy;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float mx = -INFINITY, my = -INFINITY;
  for (int64_t j = threadIdx.x; j < cols; j += blockDim.x) { mx = fmaxf(mx, x[j]); my = fmaxf(my, y[j]); }
  mx = warp_max(mx); my = warp_max(my);
  if (lane == 0) { redf[w] = mx; redf[16 + w] = my; }
  __syncthreads();
  mx = redf[0]; my = redf[16];
  for (int k = 1; k < nw; ++k) { mx = fmaxf(mx, redf[k]); my = fmaxf(my, redf[16 + k]); }
  double sx = 0.0, sy = 0.0;
  for (int64_t j = threadIdx.x; j < cols; j += blockDim.x) { sx += (double)expf(x[j] - mx); sy += (double)expf(y[j] - my); }
  sx = warp_sum_d(sx); sy = warp_sum_d(sy);
  __syncthreads();
  if (lane == 0) { redd[w] = sx; redd[16 + w] = sy; }
  __syncthreads();
  sx = 0.0; sy = 0.0;
  for (int k = 0; k < nw; ++k) { sx += redd[k]; sy += redd[16 + k]; }
  const double lx = (double)mx + log(sx), ly = (double)my + log(sy);
  double kl = 0.0;
  for (int64_t j = threadIdx.x; j < cols; j += blockDim.x) {
    const double lp = (double)x[j] - lx, lq = (double)y[j] - ly;
    kl += exp(lp) * (lp - lq);
  }
  __syncthreads();
  block_atomic_add_d(kl, out, redd);
}

}  // namespace

extern "C" {

int mcgra_noise_clamp(float* M, const float* noise, float eps, int64_t count, void* stream) {
  if (count <= 0) return 0;
  k_noise_clamp<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(M, noise, eps, count);
  MCGRA_LAUNCH_CHECK();
  return 0;
}

int mcgra_row_kl(const float* X, const float* Y, int64_t rows, int64_t cols, int64_t ldx, int64_t ldy, double* out,
                 void* stream) {
  if (rows <= 0 || cols <= 0) return 0;
  k_row_kl<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(X, Y, cols, ldx, ldy, out);
  MCGRA_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
