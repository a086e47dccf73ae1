// The n x d prior terms under the contraction measures (topology_attack.py:237-272 with --measure HSIC / CKA / DP):
//   c9  = w9  * calc(H_A[idx],  em[idx])                  (n x 16 against n x 16)
//   c10 = w10 * calc(Y_A[idx],  softmax(output2[idx]))    (n x c  against n x c; Y_A holds log-probabilities, :264-271)
// with calc = CudaCKA.linear_HSIC / linear_CKA (utils.py:1080-1091) or PGDAttack.dot_product (:480-481).  On n x d
// operands every one of them is a function of weighted second moments:
//   linear_HSIC(X, Y) = || Xc^T Yc ||_F^2,  Xc^T Yc = m (Sxy - sx sy^T)     (sx = sum_i w_i x_i, Sxy = sum_i w_i x_i y_i^T,
//                                                                            w_i = multiplicity of node i in idx / m)
//   dot_product(X, Y) = || Y^T X ||_F = m || Sxy ||_F
// so forward AND backward are O(n d d'): moments (mcgra_cross_moments), a one-block coefficient kernel, and a per-node
// gradient kernel  g_i = w_i m [ (x_i - cx) P + (y_i - cy) Q ]  chained through the second head's softmax for c10.
#include "common.cuh"

namespace {

constexpr int ND_MAXD = 32;

struct NdTerm {
  int dx, dy;
  const double* mom;      // [dx | dy | dx*dy | dy*dy] weighted moments of (X, Y)
  const double* momxx;    // same layout for (X, X)  (CKA only)
  float* coef;            // P [dx*dy] | Q [dy*dy] | cx [dx] | cy [dy]
  double weight;          // signed term weight
  int acc_slot;
};

__device__ void nd_coef_term(const NdTerm t, int measure, double m, double* acc) {
  // executed by thread 0 of a single block: sizes are <= 32 x 32
  const int dx = t.dx, dy = t.dy;
  const double* sx = t.mom;
  const double* sy = t.mom + dx;
  const double* Sxy = t.mom + dx + dy;
  const double* Syy = Sxy + dx * dy;
  float* P = t.coef;
  float* Q = P + dx * dy;
  float* cx = Q + dy * dy;
  float* cy = cx + dx;
  double value = 0.0;
  if (measure == MCGRA_M_DP) {
    double s = 0.0;
    for (int e = 0; e < dx * dy; ++e) s += Sxy[e] * Sxy[e];
    const double V = m * sqrt(s);
    value = V;
    for (int e = 0; e < dx * dy; ++e) P[e] = (float)(V > 0.0 ? t.weight * m * Sxy[e] / V : 0.0);
    for (int e = 0; e < dy * dy; ++e) Q[e] = 0.f;
    for (int k = 0; k < dx; ++k) cx[k] = 0.f;
    for (int k = 0; k < dy; ++k) cy[k] = 0.f;
  } else {
    double hxy = 0.0;
    for (int k = 0; k < dx; ++k)
      for (int l = 0; l < dy; ++l) {
        const double c = m * (Sxy[k * dy + l] - sx[k] * sy[l]);
        hxy += c * c;
      }
    double scaleP = 2.0, scaleQ = 0.0;
    value = hxy;
    if (measure == MCGRA_M_CKA) {
      double hyy = 0.0, hxx = 0.0;
      for (int k = 0; k < dy; ++k)
        for (int l = 0; l < dy; ++l) {
          const double c = m * (Syy[k * dy + l] - sy[k] * sy[l]);
          hyy += c * c;
        }
      const double* sxx = t.momxx;
      const double* Sxx = t.momxx + 2 * dx;
      for (int k = 0; k < dx; ++k)
        for (int l = 0; l < dx; ++l) {
          const double c = m * (Sxx[k * dx + l] - sxx[k] * sxx[l]);
          hxx += c * c;
        }
      const double den = sqrt(hxx) * sqrt(hyy);
      value = hxy / den;
      scaleP = 2.0 / den;
      scaleQ = -2.0 * value / hyy;
    }
    for (int k = 0; k < dx; ++k)
      for (int l = 0; l < dy; ++l) P[k * dy + l] = (float)(t.weight * scaleP * m * (Sxy[k * dy + l] - sx[k] * sy[l]));
    for (int k = 0; k < dy; ++k)
      for (int l = 0; l < dy; ++l) Q[k * dy + l] = (float)(t.weight * scaleQ * m * (Syy[k * dy + l] - sy[k] * sy[l]));
    for (int k = 0; k < dx; ++k) cx[k] = (float)sx[k];
    for (int k = 0; k < dy; ++k) cy[k] = (float)sy[k];
  }
  acc[t.acc_slot] += t.weight * value;
}

__global__ void k_nd_coef(NdTerm t9, NdTerm t10, int on9, int on10, int measure, double m, double* acc) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (on9) nd_coef_term(t9, measure, m, acc);
  if (on10) nd_coef_term(t10, measure, m, acc);
}

// p2 = softmax(em Wl^T + bl) (the second head's probabilities, :259-271)
__global__ void k_nd_p2(const float* __restrict__ em, const float* __restrict__ Wl, const float* __restrict__ bl, int64_t n, int c,
                        float* __restrict__ p2) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float e[HID], z[ND_MAXD];
#pragma unroll
  for (int k = 0; k < HID; ++k) e[k] = em[i * HID + k];
  float mx = -INFINITY;
  for (int q = 0; q < c; ++q) {
    float s = bl[q];
#pragma unroll
    for (int k = 0; k < HID; ++k) s = fmaf(e[k], Wl[q * HID + k], s);
    z[q] = s;
    mx = fmaxf(mx, s);
  }
  float den = 0.f;
  for (int q = 0; q < c; ++q) { z[q] = expf(z[q] - mx); den += z[q]; }
  for (int q = 0; q < c; ++q) p2[i * c + q] = z[q] / den;
}

__global__ void k_nd_grad(const float* __restrict__ HA, const float* __restrict__ em, const float* __restrict__ YA,
                          const float* __restrict__ p2, const float* __restrict__ Wl, const float* __restrict__ wmult,
                          int64_t n, int c, double m, const float* __restrict__ coef9, const float* __restrict__ coef10,
                          int on9, int on10, float* __restrict__ demd) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float wi = wmult[i] * (float)m;
  if (wi == 0.f) return;
  float g[HID];
#pragma unroll
  for (int k = 0; k < HID; ++k) g[k] = 0.f;
  if (on9) {
    const float* P = coef9;
    const float* Q = P + HID * HID;
    const float* cx = Q + HID * HID;
    const float* cy = cx + HID;
    for (int k = 0; k < HID; ++k) {
      const float xk = HA[i * HID + k] - cx[k], yk = em[i * HID + k] - cy[k];
#pragma unroll
      for (int l = 0; l < HID; ++l) g[l] = fmaf(xk, P[k * HID + l], fmaf(yk, Q[k * HID + l], g[l]));
    }
  }
  if (on10) {
    const float* P = coef10;
    const float* Q = P + c * c;
    const float* cx = Q + c * c;
    const float* cy = cx + c;
    float gp[ND_MAXD], p[ND_MAXD];
    for (int l = 0; l < c; ++l) { gp[l] = 0.f; p[l] = p2[i * c + l]; }
    for (int k = 0; k < c; ++k) {
      const float xk = YA[i * c + k] - cx[k], yk = p[k] - cy[k];
      for (int l = 0; l < c; ++l) gp[l] = fmaf(xk, P[k * c + l], fmaf(yk, Q[k * c + l], gp[l]));
    }
    float dotp = 0.f;                                          // softmax backward: dz = p o (gp - <gp, p>)
    for (int l = 0; l < c; ++l) dotp = fmaf(gp[l], p[l], dotp);
    for (int l = 0; l < c; ++l) {
      const float dz = p[l] * (gp[l] - dotp);
#pragma unroll
      for (int k = 0; k < HID; ++k) g[k] = fmaf(dz, Wl[l * HID + k], g[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < HID; ++k) demd[i * HID + k] += wi * g[k];
}

}  // namespace

extern "C" {

int mcgra_cross_moments(const float* X, int dx, const float* Y, int dy, const float* w, int64_t n, double* out, void* stream);

int64_t mcgra_nd_scratch_doubles(int nclass) {
  const int d = HID, c = nclass;
  return 2 * (2 * d + 2 * d * d) + 2 * (2 * c + 2 * c * c);
}
int64_t mcgra_nd_scratch_floats(int nclass) {
  const int d = HID, c = nclass;
  return (2 * d * d + 2 * d) + (2 * c * c + 2 * c);
}

int mcgra_nd_measure(const mcgra_nd_args* a, void* stream) {
  if (a == nullptr) return -1;
  const int c = a->nclass, d = HID;
  if (c < 1 || c > ND_MAXD) return -2;
  if (a->measure != MCGRA_M_HSIC && a->measure != MCGRA_M_CKA && a->measure != MCGRA_M_DP) return -3;
  const int on9 = a->w9 != 0.f, on10 = a->w10 != 0.f;
  if (!on9 && !on10) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = a->n;
  const int64_t nd = mcgra_nd_scratch_doubles(c);
  cudaError_t e = cudaMemsetAsync(a->mom, 0, sizeof(double) * (size_t)nd, st);
  if (e != cudaSuccess) return (int)e;
  double* mom9 = a->mom;
  double* mom9xx = mom9 + (2 * d + 2 * d * d);
  double* mom10 = mom9xx + (2 * d + 2 * d * d);
  double* mom10xx = mom10 + (2 * c + 2 * c * c);
  float* coef9 = a->coef;
  float* coef10 = coef9 + (2 * d * d + 2 * d);
  const bool cka = a->measure == MCGRA_M_CKA;
  int rc;
  if (on9) {
    if ((rc = mcgra_cross_moments(a->HA, d, a->em, d, a->wmult, n, mom9, stream)) != 0) return rc;
    if (cka && (rc = mcgra_cross_moments(a->HA, d, a->HA, d, a->wmult, n, mom9xx, stream)) != 0) return rc;
  }
  if (on10) {
    k_nd_p2<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(a->em, a->Wl, a->bl, n, c, a->p2);
    if ((rc = mcgra_cross_moments(a->YA, c, a->p2, c, a->wmult, n, mom10, stream)) != 0) return rc;
    if (cka && (rc = mcgra_cross_moments(a->YA, c, a->YA, c, a->wmult, n, mom10xx, stream)) != 0) return rc;
  }
  NdTerm t9 = {d, d, mom9, mom9xx, coef9, (double)a->w9, MCGRA_ACC_C9};
  NdTerm t10 = {c, c, mom10, mom10xx, coef10, (double)a->w10, MCGRA_ACC_C10};
  k_nd_coef<<<1, 32, 0, st>>>(t9, t10, on9, on10, a->measure, a->m, a->acc);
  k_nd_grad<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(a->HA, a->em, a->YA, a->p2, a->Wl, a->wmult, n, c, a->m, coef9, coef10,
                                                         on9, on10, a->demd);
  MCGRA_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
