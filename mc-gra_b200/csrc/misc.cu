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

// ---- feature smoothing of the GraphMI attack (MC-GPB/topology_attack.py:57-61, feature_smoothing :163-177) ----------
//   S = tr(X^T L~ X) = sum_i rt_i^2 d_i G_ii - sum_{i != j} rt_i rt_j M_ij G_ij,   G = X X^T, d = M 1, rt = (d + 1e-3)^-1/2
//   dS/dx_(ij) = -2 rt_i rt_j G_ij + rho_i + rho_j,   rho_i = dS/dd_i = 1e-3 G_ii rt_i^4 + rt_i^3 t_i,  t_i = sum_j rt_j M_ij G_ij
// One pass over the x and G tiles writes the element-wise part (scaled by coef) as gradient tiles for mcgra_fold_adam
// (Gtiles) and accumulates t; the node kernel turns t into rho and the value.
__global__ void k_smooth_rt(int64_t n, const float* __restrict__ d1, float* __restrict__ rt, float* __restrict__ trow) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float d = d1[i] - 1.f;                    // the engine's degree carries the +1 of normalize_adj_tensor
  const float v = rsqrtf(d + 1e-3f);
  rt[i] = isinf(v) ? 0.f : v;
  trow[i] = 0.f;
}
__global__ void __launch_bounds__(256)
k_smooth_pass(const float* __restrict__ tiles, const float* __restrict__ Gfeat, int64_t n, int64_t t0, const float* mu, int raw,
              const float* __restrict__ rt, float coef, float* __restrict__ Gt, float* __restrict__ trow) {
  __shared__ float rI[TILE], rJ[TILE], col[2][TILE];
  int I, J;
  tile_coords(t0 + blockIdx.x, I, J);
  const ParamView pv = load_view(mu, raw);
  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t i0 = (int64_t)I * TILE, j0 = (int64_t)J * TILE;
  if (tid < TILE) {
    rI[tid] = i0 + tid < n ? rt[i0 + tid] : 0.f;
    rJ[tid] = j0 + tid < n ? rt[j0 + tid] : 0.f;
  }
  __syncthreads();
  const float* xs = tiles + (int64_t)blockIdx.x * TILE_ELEMS;
  const float* gs = Gfeat + (int64_t)blockIdx.x * TILE_ELEMS;
  float* out = Gt + (int64_t)blockIdx.x * TILE_ELEMS;
  const int b = tid & 127, ah = tid >> 7;
  float c = 0.f;
  for (int it = 0; it < TILE / 2; ++it) {
    const int row = 2 * it + ah;
    const int64_t gi = i0 + row, gj = j0 + b;
    float o = 0.f, rsum = 0.f;
    if (gj < gi && gi < n) {
      const float g = gs[row * TILE + b], M = pv.adj(xs[row * TILE + b]);
      o = -2.f * coef * rI[row] * rJ[b] * g;
      rsum = rJ[b] * M * g;
      c += rI[row] * M * g;
    }
    out[row * TILE + b] = o;
    rsum = warp_sum(rsum);
    if (lane == 0 && gi < n && rsum != 0.f) atomicAdd(trow + gi, rsum);
  }
  col[ah][b] = c;
  __syncthreads();
  if (tid < TILE && j0 + tid < n) {
    const float s = col[0][tid] + col[1][tid];
    if (s != 0.f) atomicAdd(trow + j0 + tid, s);
  }
}
__global__ void k_smooth_node(int64_t n, const float* __restrict__ d1, const float* __restrict__ rt, const float* __restrict__ trow,
                              const float* __restrict__ gdiag, float coef, float* __restrict__ rho, double* __restrict__ acc_slot) {
  __shared__ double red[32];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0.0;
  if (i < n) {
    const double d = (double)d1[i] - 1.0, r = (double)rt[i], t = (double)trow[i], q = (double)gdiag[i];
    v = r * r * d * q - r * t;
    rho[i] += (float)((double)coef * (1e-3 * q * r * r * r * r + r * r * r * t));
  }
  block_atomic_add_d(v * (double)coef, acc_slot, red);
}

}  // namespace

extern "C" {

int mcgra_smooth(const float* tiles, const float* Gfeat, const float* gdiag, int64_t n, int tr0, int tr1, const float* mu,
                 int raw, const float* d, float coef, float* rt, float* trow, float* Gt, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nt = tri(tr1) - tri(tr0);
  k_smooth_rt<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, d, rt, trow);
  if (nt > 0) k_smooth_pass<<<(unsigned)nt, 256, 0, st>>>(tiles, Gfeat, n, tri(tr0), mu, raw, rt, coef, Gt, trow);
  MCGRA_LAUNCH_CHECK();
  (void)gdiag;
  return 0;
}

int mcgra_smooth_node(int64_t n, const float* d, const float* rt, const float* trow, const float* gdiag, float coef,
                      float* rho, double* acc_slot, void* stream) {
  k_smooth_node<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, d, rt, trow, gdiag, coef, rho, acc_slot);
  MCGRA_LAUNCH_CHECK();
  return 0;
}

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
