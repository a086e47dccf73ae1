// KL measure on two n x n operands (PGDAttack.calc_kl, topology_attack.py:483-487, called at :212-229 with
// --measure KL): c2 = KLDiv(log_softmax(M1, 1), softmax(A_hat, 1), batchmean) and -- when it runs together with c2 --
// c1 = KLDiv(log_softmax(A_hat, 1), softmax(F, 1), batchmean), as three passes over the tiled triangle.  The gram
// M1 = relu(zhat zhat^T) is regenerated per tile from the n x 16 factors; no n x n matrix is materialised.
//
//   pass 0  seA[i] += exp(A_ij),  seM[i] += exp(M1_ij)                         (both orientations of a stored entry)
//   node 0  lseA = log(seA + exp(r_i^2)),  lseM = log(seM + 1)                 (diagonals: A_ii = r_i^2, M1_ii = 0)
//   pass 1  kl[i] += pA_ij (A_ij - M1_ij),  c1[i] += pF_ij (F_ij - A_ij)       pA = exp(A - lseA_i), pF = exp(F - lseF_i)
//   node 1  KL_i = kl[i] - lseA_i + lseM_i (fp64),  values into acc, diagonal gradient Fdiag
//   pass 2  EA_ij + EA_ji -> EAt,   EA_ij = (k2/n) pA_ij (A_ij - M1_ij - kl_i) + (k1/n) (pA_ij - pF_ij)
//           CM_ij + CM_ji -> Ct,    CM_ij = (k2/n) (qM_ij - pA_ij),            qM = exp(M1 - lseM_i)
// (d/dA of sum_j pA (log pA - log qM) through the softmax is pA (log pA - log qM - KL_i); the row constants
//  lseA_i - lseM_i cancel against KL_i, leaving A_ij - M1_ij - kl_i: no cancellation in fp32.)
// The tiles EAt / Ct / Fdiag feed the MCGRA_M_PRE path of mcgra_pairs / mcgra_fold_adam.
#include "common.cuh"

namespace {

struct Kl2Args {
  const float* tiles;     // x' shard
  const float* Ftiles;    // feature_adj shard or NULL
  const float* zhat;
  const float* r;
  const float* lseA;      // [n] (passes 1, 2)
  const float* lseM;
  const float* lseF;
  const float* klrow;     // [n] float (pass 2)
  float* acc0;            // pass 0: seA | pass 1: klrow accum
  float* acc1;            // pass 0: seM | pass 1: c1row accum
  float* EAt;
  float* Ct;
  float k1n, k2n;         // k1c / n, k2c / n
  int64_t n;
  int64_t t0;
  const float* mu;
  int raw;
};

template <int PASS>
__global__ void __launch_bounds__(256)
k_kl2_pass(const Kl2Args a) {
  __shared__ float zI[TILE][HID + 1], zJ[TILE][HID + 1];
  __shared__ float rI[TILE], rJ[TILE], sI[3][TILE], sJ[3][TILE], kI[TILE], kJ[TILE];
  __shared__ float col0[2][TILE], col1[2][TILE];
  int I, J;
  tile_coords(a.t0 + blockIdx.x, I, J);
  const ParamView pv = load_view(a.mu, a.raw);
  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t i0 = (int64_t)I * TILE, j0 = (int64_t)J * TILE, n = a.n;
  const float* src = a.tiles + (int64_t)blockIdx.x * TILE_ELEMS;
  const float* fsrc = a.Ftiles != nullptr ? a.Ftiles + (int64_t)blockIdx.x * TILE_ELEMS : nullptr;
  for (int e = tid; e < TILE * HID; e += 256) {
    const int row = e / HID, k = e % HID;
    zI[row][k] = i0 + row < n ? a.zhat[(i0 + row) * HID + k] : 0.f;
    zJ[row][k] = j0 + row < n ? a.zhat[(j0 + row) * HID + k] : 0.f;
  }
  if (tid < TILE) {
    const int64_t gi = i0 + tid, gj = j0 + tid;
    rI[tid] = gi < n ? a.r[gi] : 0.f;
    rJ[tid] = gj < n ? a.r[gj] : 0.f;
    if (PASS >= 1) {
      sI[0][tid] = gi < n ? a.lseA[gi] : 0.f;
      sJ[0][tid] = gj < n ? a.lseA[gj] : 0.f;
      sI[1][tid] = gi < n ? a.lseM[gi] : 0.f;
      sJ[1][tid] = gj < n ? a.lseM[gj] : 0.f;
      sI[2][tid] = (gi < n && a.lseF != nullptr) ? a.lseF[gi] : 0.f;
      sJ[2][tid] = (gj < n && a.lseF != nullptr) ? a.lseF[gj] : 0.f;
    }
    if (PASS == 2) {
      kI[tid] = gi < n ? a.klrow[gi] : 0.f;
      kJ[tid] = gj < n ? a.klrow[gj] : 0.f;
    }
  }
  __syncthreads();
  const int b = tid & 127, ah = tid >> 7;            // column of the tile, row parity
  float c0 = 0.f, c1 = 0.f;                          // column (mirrored) accumulators of this thread's column b
  for (int it = 0; it < TILE / 2; ++it) {
    const int row = 2 * it + ah;
    const int64_t gi = i0 + row, gj = j0 + b;
    const bool valid = gj < gi && gi < n;
    float r0 = 0.f, r1 = 0.f;
    float ea = 0.f, cm = 0.f;
    if (valid) {
      const float A = rI[row] * pv.adj(src[row * TILE + b]) * rJ[b];
      float d = 0.f;
#pragma unroll
      for (int k = 0; k < HID; ++k) d = fmaf(zI[row][k], zJ[b][k], d);
      const float M = fmaxf(d, 0.f);
      if (PASS == 0) {
        const float eA = __expf(A), eM = __expf(M);
        r0 = eA; c0 += eA;
        r1 = eM; c1 += eM;
      } else {
        const float pAi = __expf(A - sI[0][row]), pAj = __expf(A - sJ[0][b]);
        const float Fv = fsrc != nullptr ? fsrc[row * TILE + b] : 0.f;
        const float pFi = fsrc != nullptr ? __expf(Fv - sI[2][row]) : 0.f;
        const float pFj = fsrc != nullptr ? __expf(Fv - sJ[2][b]) : 0.f;
        if (PASS == 1) {
          r0 = pAi * (A - M); c0 += pAj * (A - M);
          r1 = pFi * (Fv - A); c1 += pFj * (Fv - A);
        } else {
          const float qMi = __expf(M - sI[1][row]), qMj = __expf(M - sJ[1][b]);
          ea = a.k2n * (pAi * (A - M - kI[row]) + pAj * (A - M - kJ[b])) + a.k1n * ((pAi - pFi) + (pAj - pFj));
          cm = a.k2n * ((qMi - pAi) + (qMj - pAj));
        }
      }
    }
    if (PASS == 2) {
      a.EAt[(int64_t)blockIdx.x * TILE_ELEMS + row * TILE + b] = ea;
      a.Ct[(int64_t)blockIdx.x * TILE_ELEMS + row * TILE + b] = cm;
    } else {
      r0 = warp_sum(r0);
      r1 = warp_sum(r1);
      if (lane == 0 && gi < n) {
        if (r0 != 0.f) atomicAdd(a.acc0 + gi, r0);
        if (r1 != 0.f) atomicAdd(a.acc1 + gi, r1);
      }
    }
  }
  if (PASS != 2) {
    col0[ah][b] = c0;
    col1[ah][b] = c1;
    __syncthreads();
    if (tid < TILE && j0 + tid < n) {
      const float s0 = col0[0][tid] + col0[1][tid], s1 = col1[0][tid] + col1[1][tid];
      if (s0 != 0.f) atomicAdd(a.acc0 + j0 + tid, s0);
      if (s1 != 0.f) atomicAdd(a.acc1 + j0 + tid, s1);
    }
  }
}

// node 0: lse from the all-reduced exp sums; node 1: KL_i, loss values, diagonal gradient
__global__ void k_kl2_node0(int64_t n, const float* __restrict__ r, const float* __restrict__ seA,
                            const float* __restrict__ seM, float* __restrict__ lseA, float* __restrict__ lseM) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double ri = (double)r[i];
  lseA[i] = (float)log((double)seA[i] + exp(ri * ri));
  lseM[i] = (float)log((double)seM[i] + 1.0);
}
__global__ void k_kl2_node1(int64_t n, const float* __restrict__ r, const float* __restrict__ seA,
                            const float* __restrict__ seM, const float* __restrict__ lseF, const float* __restrict__ Fdiag_feat,
                            float* __restrict__ klrow, const float* __restrict__ c1row, double k1c, double k2c,
                            float* __restrict__ Fdiag, double* __restrict__ acc) {
  __shared__ double red[32];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double v1 = 0.0, v2 = 0.0;
  if (i < n) {
    const double ri = (double)r[i], Aii = ri * ri;
    const double lA = log((double)seA[i] + exp(Aii)), lM = log((double)seM[i] + 1.0);
    const double pAii = exp(Aii - lA);
    double g = 0.0;
    if (k2c != 0.0) {
      const double kl = (double)klrow[i] + pAii * (Aii - 0.0);          // diagonal entry: M1_ii = 0
      klrow[i] = (float)kl;
      v2 = kl - lA + lM;
      g += (k2c / (double)n) * pAii * (Aii - kl);
    }
    if (k1c != 0.0) {
      const double Fii = (double)Fdiag_feat[i], lF = (double)lseF[i];
      const double pFii = exp(Fii - lF);
      v1 = (double)c1row[i] + pFii * (Fii - Aii) - (lF - lA);
      g += (k1c / (double)n) * (pAii - pFii);
    }
    Fdiag[i] = (float)g;
  }
  block_atomic_add_d(v1 * k1c / (double)n, acc + MCGRA_ACC_C1D, red);
  block_atomic_add_d(v2 * k2c / (double)n, acc + MCGRA_ACC_C2D, red);
}

}  // namespace

extern "C" {

int mcgra_kl2_pass(int pass, const mcgra_kl2_args* k, void* stream) {
  if (k == nullptr || pass < 0 || pass > 2) return -1;
  const int64_t nt = tri(k->tr1) - tri(k->tr0);
  if (nt <= 0) return 0;
  Kl2Args a;
  a.tiles = k->tiles; a.Ftiles = k->Ftiles; a.zhat = k->zhat; a.r = k->r;
  a.lseA = k->lseA; a.lseM = k->lseM; a.lseF = k->lseF; a.klrow = k->klrow;
  a.acc0 = pass == 0 ? k->seA : k->klrow;
  a.acc1 = pass == 0 ? k->seM : k->c1row;
  a.EAt = k->EAt; a.Ct = k->Ct;
  a.k1n = (float)(k->k1c / (double)k->n); a.k2n = (float)(k->k2c / (double)k->n);
  a.n = k->n; a.t0 = tri(k->tr0); a.mu = k->mu; a.raw = k->raw;
  cudaStream_t st = (cudaStream_t)stream;
  if (pass == 0) k_kl2_pass<0><<<(unsigned)nt, 256, 0, st>>>(a);
  else if (pass == 1) k_kl2_pass<1><<<(unsigned)nt, 256, 0, st>>>(a);
  else k_kl2_pass<2><<<(unsigned)nt, 256, 0, st>>>(a);
  MCGRA_LAUNCH_CHECK();
  return 0;
}

int mcgra_kl2_node(int stage, const mcgra_kl2_args* k, float* Fdiag, double* acc, void* stream) {
  if (k == nullptr) return -1;
  const unsigned g = (unsigned)((k->n + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (stage == 0) k_kl2_node0<<<g, 256, 0, st>>>(k->n, k->r, k->seA, k->seM, k->lseA, k->lseM);
  else k_kl2_node1<<<g, 256, 0, st>>>(k->n, k->r, k->seA, k->seM, k->lseF, k->Fdiag_feat, k->klrow, k->c1row, k->k1c, k->k2c,
                                      Fdiag, acc);
  MCGRA_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
