// Propagation over the tiled symmetric adjacency estimate: degree, Y += M*B, row sum-exp.
//
// Reference work replaced: GraphConvolution.forward's torch.spmm(adj, support) (MC-GRA/models/gcn.py:42) and its
// transposed product in autograd, the row sums of utils.normalize_adj_tensor (MC-GRA/utils.py:224-225), and the
// element-wise loss terms c1 (MSE/KL vs feature_adj) and c6 (Info_entropy) over A_hat
// (MC-GRA/topology_attack.py:212-232, 44-47) which are fused into the first propagation pass.
//
// Tile-symmetric schedule: one CTA per 128x128 tile (I,J) of the lower triangle.  The tile is read from HBM once
// and used twice:  Y[I rows] += T * B[J rows]  and  Y[J rows] += T^T * B[I rows].  The unnormalised / normalised
// n x n matrices are never materialised: the D^-1/2 scaling lives in the B operand (prologue, node kernels) and in
// the consumers of Y (epilogue, node kernels).
//
// v1 engine: fp32 FFMA with a 2 x K register tile per thread (exact fp32 accumulate).  HBM traffic = one read of
// the tile shard (4 bytes per stored entry) + O(n K).
#include "common.cuh"

namespace {

constexpr int XS_LD = TILE + 4;   // padded row stride (floats): LDS.128 along a row is conflict free

template <int KC>
struct PropSmem {
  float xs[TILE][XS_LD];
  float bj[TILE][KC];
  float bi[TILE][KC];
  float rI[TILE], rJ[TILE], lseAI[TILE], lseAJ[TILE], lseFI[TILE], lseFJ[TILE], dlI[TILE], dlJ[TILE];
  float rowacc[TILE], colacc[TILE];
  double red[32];
};

template <int KC, bool ELEM>
__global__ void __launch_bounds__(128, 2)
k_propagate(const float* __restrict__ tiles, int64_t n, int64_t t0, const float* mu, int raw,
            const float* __restrict__ B, float* __restrict__ Y, mcgra_elem_args ea) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PropSmem<KC>& sm = *reinterpret_cast<PropSmem<KC>*>(smem_raw);
  int I, J;
  tile_coords(t0 + blockIdx.x, I, J);
  const ParamView pv = load_view(mu, raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t i0 = (int64_t)I * TILE, j0 = (int64_t)J * TILE;
  const float4* src = reinterpret_cast<const float4*>(tiles + (int64_t)blockIdx.x * TILE_ELEMS);

  // ---- stage B rows of both sides -------------------------------------------------------------------------
  for (int e = tid; e < TILE * KC / 4; e += 128) {
    const int row = e / (KC / 4), c4 = e % (KC / 4);
    float4 vj = make_float4(0.f, 0.f, 0.f, 0.f), vi = vj;
    if (j0 + row < n) vj = reinterpret_cast<const float4*>(B + (j0 + row) * KC)[c4];
    if (i0 + row < n) vi = reinterpret_cast<const float4*>(B + (i0 + row) * KC)[c4];
    reinterpret_cast<float4*>(&sm.bj[row][0])[c4] = vj;
    reinterpret_cast<float4*>(&sm.bi[row][0])[c4] = vi;
  }
  if (ELEM) {
    const int64_t gi = i0 + tid, gj = j0 + tid;
    sm.rI[tid] = gi < n ? ea.r[gi] : 0.f;
    sm.rJ[tid] = gj < n ? ea.r[gj] : 0.f;
    if (ea.measure == MCGRA_M_KL) {
      sm.lseAI[tid] = gi < n ? ea.lseA[gi] : 0.f;
      sm.lseAJ[tid] = gj < n ? ea.lseA[gj] : 0.f;
      sm.lseFI[tid] = gi < n ? ea.lseF[gi] : 0.f;
      sm.lseFJ[tid] = gj < n ? ea.lseF[gj] : 0.f;
      sm.dlI[tid] = gi < n ? ea.dlse[gi] : 0.f;
      sm.dlJ[tid] = gj < n ? ea.dlse[gj] : 0.f;
    }
    sm.colacc[tid] = 0.f;
    __syncthreads();
  }

  // ---- stage the tile (one warp = one 512 B row per step), fused element-wise terms -----------------------
  float col_e[4] = {0.f, 0.f, 0.f, 0.f};
  float v1 = 0.f, v6 = 0.f;
  const float4* fsrc = (ELEM && ea.Ftiles != nullptr)
                           ? reinterpret_cast<const float4*>(ea.Ftiles + (int64_t)blockIdx.x * TILE_ELEMS)
                           : nullptr;
#pragma unroll 4
  for (int it = 0; it < 32; ++it) {
    const int row = it * 4 + warp;
    const int idx = row * 32 + lane;
    const float4 raw4 = src[idx];
    const int64_t gi = i0 + row;
    const int64_t gj = j0 + lane * 4;
    float xv[4] = {raw4.x, raw4.y, raw4.z, raw4.w};
    bool ok[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      ok[k] = (gj + k < gi) && (gi < n);
      xv[k] = ok[k] ? pv.adj(xv[k]) : 0.f;
    }
    *reinterpret_cast<float4*>(&sm.xs[row][lane * 4]) = make_float4(xv[0], xv[1], xv[2], xv[3]);
    if (ELEM) {
      float4 f4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (fsrc != nullptr) f4 = fsrc[idx];
      const float fv[4] = {f4.x, f4.y, f4.z, f4.w};
      const float ri = sm.rI[row];
      float row_e = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (!ok[k]) continue;
        const float rj = sm.rJ[lane * 4 + k];
        const float ah = ri * xv[k] * rj;
        float esym = 0.f;   // e'_ij + e'_ji
        if (ea.measure == MCGRA_M_MSE) {
          const float df = ah - fv[k];
          v1 += 2.f * df * df;
          esym += 4.f * ea.k1 * df;
        } else if (ea.measure == MCGRA_M_KL) {
          const float xij = __expf(fv[k] - sm.lseFI[row]);
          const float xji = __expf(fv[k] - sm.lseFJ[lane * 4 + k]);
          const float lij = ah - sm.lseAI[row];
          const float lji = ah - sm.lseAJ[lane * 4 + k];
          v1 += xij * ((fv[k] - ah) - sm.dlI[row]) + xji * ((fv[k] - ah) - sm.dlJ[lane * 4 + k]);
          esym += ea.k1 * ((__expf(lij) - xij) + (__expf(lji) - xji));
        }
        if (ea.k6 != 0.f) {
          v6 += 2.f * ent_val(ah);
          esym += 2.f * ea.k6 * ent_grad(ah);
        }
        const float t = esym * xv[k];
        row_e += t * rj;
        col_e[k] += t * ri;
      }
      row_e = warp_sum(row_e);
      if (lane == 0) sm.rowacc[row] = row_e;
    }
  }
  if (ELEM) {
#pragma unroll
    for (int k = 0; k < 4; ++k) atomicAdd(&sm.colacc[lane * 4 + k], col_e[k]);
  }
  __syncthreads();

  // ---- compute: warps 0-1 direct product (rows of I), warps 2-3 mirrored product (rows of J) -------------
  const int u = tid & 63;
  float acc0[KC], acc1[KC];
#pragma unroll
  for (int c = 0; c < KC; ++c) acc0[c] = acc1[c] = 0.f;

  if (warp < 2) {
    // Y[i0+u], Y[i0+u+64] += sum_b xs[u][b] * bj[b][:]
#pragma unroll 1
    for (int b = 0; b < TILE; b += 4) {
      const float4 xa = *reinterpret_cast<const float4*>(&sm.xs[u][b]);
      const float4 xb = *reinterpret_cast<const float4*>(&sm.xs[u + 64][b]);
      const float xa_[4] = {xa.x, xa.y, xa.z, xa.w};
      const float xb_[4] = {xb.x, xb.y, xb.z, xb.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int c4 = 0; c4 < KC / 4; ++c4) {
          const float4 bv = *reinterpret_cast<const float4*>(&sm.bj[b + k][c4 * 4]);
          acc0[c4 * 4 + 0] = fmaf(xa_[k], bv.x, acc0[c4 * 4 + 0]);
          acc0[c4 * 4 + 1] = fmaf(xa_[k], bv.y, acc0[c4 * 4 + 1]);
          acc0[c4 * 4 + 2] = fmaf(xa_[k], bv.z, acc0[c4 * 4 + 2]);
          acc0[c4 * 4 + 3] = fmaf(xa_[k], bv.w, acc0[c4 * 4 + 3]);
          acc1[c4 * 4 + 0] = fmaf(xb_[k], bv.x, acc1[c4 * 4 + 0]);
          acc1[c4 * 4 + 1] = fmaf(xb_[k], bv.y, acc1[c4 * 4 + 1]);
          acc1[c4 * 4 + 2] = fmaf(xb_[k], bv.z, acc1[c4 * 4 + 2]);
          acc1[c4 * 4 + 3] = fmaf(xb_[k], bv.w, acc1[c4 * 4 + 3]);
        }
      }
    }
  } else {
    // Y[j0+u], Y[j0+u+64] += sum_a xs[a][u] * bi[a][:]
#pragma unroll 1
    for (int a = 0; a < TILE; a += 4) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float xa = sm.xs[a + k][u];
        const float xb = sm.xs[a + k][u + 64];
#pragma unroll
        for (int c4 = 0; c4 < KC / 4; ++c4) {
          const float4 bv = *reinterpret_cast<const float4*>(&sm.bi[a + k][c4 * 4]);
          acc0[c4 * 4 + 0] = fmaf(xa, bv.x, acc0[c4 * 4 + 0]);
          acc0[c4 * 4 + 1] = fmaf(xa, bv.y, acc0[c4 * 4 + 1]);
          acc0[c4 * 4 + 2] = fmaf(xa, bv.z, acc0[c4 * 4 + 2]);
          acc0[c4 * 4 + 3] = fmaf(xa, bv.w, acc0[c4 * 4 + 3]);
          acc1[c4 * 4 + 0] = fmaf(xb, bv.x, acc1[c4 * 4 + 0]);
          acc1[c4 * 4 + 1] = fmaf(xb, bv.y, acc1[c4 * 4 + 1]);
          acc1[c4 * 4 + 2] = fmaf(xb, bv.z, acc1[c4 * 4 + 2]);
          acc1[c4 * 4 + 3] = fmaf(xb, bv.w, acc1[c4 * 4 + 3]);
        }
      }
    }
  }

  // ---- flush: vector reductions into Y ---------------------------------------------------------------------
  {
    const int64_t base = (warp < 2) ? i0 : j0;
    const int64_t r0 = base + u, r1 = base + u + 64;
    if (r0 < n) {
      float4* dst = reinterpret_cast<float4*>(Y + r0 * KC);
#pragma unroll
      for (int c4 = 0; c4 < KC / 4; ++c4)
        atomicAdd(dst + c4, make_float4(acc0[c4 * 4], acc0[c4 * 4 + 1], acc0[c4 * 4 + 2], acc0[c4 * 4 + 3]));
    }
    if (r1 < n) {
      float4* dst = reinterpret_cast<float4*>(Y + r1 * KC);
#pragma unroll
      for (int c4 = 0; c4 < KC / 4; ++c4)
        atomicAdd(dst + c4, make_float4(acc1[c4 * 4], acc1[c4 * 4 + 1], acc1[c4 * 4 + 2], acc1[c4 * 4 + 3]));
    }
  }
  if (ELEM) {
    const int64_t gi = i0 + tid, gj = j0 + tid;
    if (gi < n && sm.rowacc[tid] != 0.f) atomicAdd(ea.eps_row + gi, sm.rowacc[tid]);
    if (gj < n && sm.colacc[tid] != 0.f) atomicAdd(ea.eps_row + gj, sm.colacc[tid]);
    if (ea.measure != MCGRA_M_NONE) block_atomic_add_d((double)v1 * (double)ea.k1, ea.acc + MCGRA_ACC_C1, sm.red);
    if (ea.k6 != 0.f) block_atomic_add_d((double)v6 * (double)ea.k6, ea.acc + MCGRA_ACC_C6, sm.red);
  }
}

// degree: one CTA per tile, 128 threads, row sums through warp shuffles, column sums in registers
__global__ void __launch_bounds__(128)
k_degree(const float* __restrict__ tiles, int64_t n, int64_t t0, const float* mu, int raw, float* __restrict__ d) {
  __shared__ float colacc[TILE];
  int I, J;
  tile_coords(t0 + blockIdx.x, I, J);
  const ParamView pv = load_view(mu, raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t i0 = (int64_t)I * TILE, j0 = (int64_t)J * TILE;
  const float4* src = reinterpret_cast<const float4*>(tiles + (int64_t)blockIdx.x * TILE_ELEMS);
  colacc[tid] = 0.f;
  __syncthreads();
  float col[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
  for (int it = 0; it < 32; ++it) {
    const int row = it * 4 + warp;
    const float4 raw4 = src[row * 32 + lane];
    const int64_t gi = i0 + row, gj = j0 + lane * 4;
    const float xv[4] = {raw4.x, raw4.y, raw4.z, raw4.w};
    float rs = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float a = ((gj + k < gi) && (gi < n)) ? pv.adj(xv[k]) : 0.f;
      rs += a;
      col[k] += a;
    }
    rs = warp_sum(rs);
    if (lane == 0 && gi < n && rs != 0.f) atomicAdd(d + gi, rs);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) atomicAdd(&colacc[lane * 4 + k], col[k]);
  __syncthreads();
  if (j0 + tid < n && colacc[tid] != 0.f) atomicAdd(d + j0 + tid, colacc[tid]);
}

// sumexp[i] += sum_{j != i, valid} exp(r_i M_ij r_j)   (both orientations of every stored entry)
__global__ void __launch_bounds__(128)
k_row_sumexp(const float* __restrict__ tiles, int64_t n, int64_t t0, const float* mu, int raw,
             const float* __restrict__ r, float* __restrict__ sumexp) {
  __shared__ float colacc[TILE], rI[TILE], rJ[TILE];
  int I, J;
  tile_coords(t0 + blockIdx.x, I, J);
  const ParamView pv = load_view(mu, raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t i0 = (int64_t)I * TILE, j0 = (int64_t)J * TILE;
  const float4* src = reinterpret_cast<const float4*>(tiles + (int64_t)blockIdx.x * TILE_ELEMS);
  colacc[tid] = 0.f;
  rI[tid] = i0 + tid < n ? r[i0 + tid] : 0.f;
  rJ[tid] = j0 + tid < n ? r[j0 + tid] : 0.f;
  __syncthreads();
  float col[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (int it = 0; it < 32; ++it) {
    const int row = it * 4 + warp;
    const float4 raw4 = src[row * 32 + lane];
    const int64_t gi = i0 + row, gj = j0 + lane * 4;
    const float xv[4] = {raw4.x, raw4.y, raw4.z, raw4.w};
    float rs = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if ((gj + k < gi) && (gi < n)) {
        const float e = __expf(rI[row] * pv.adj(xv[k]) * rJ[lane * 4 + k]);
        rs += e;
        col[k] += e;
      }
    }
    rs = warp_sum(rs);
    if (lane == 0 && gi < n && rs != 0.f) atomicAdd(sumexp + gi, rs);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) atomicAdd(&colacc[lane * 4 + k], col[k]);
  __syncthreads();
  if (j0 + tid < n && colacc[tid] != 0.f) atomicAdd(sumexp + j0 + tid, colacc[tid]);
}


// ---------------------------------------------------------------------------------------------------------
// v2 engine: warp-level tensor-core MMA (mma.sync.m16n8k8 tf32) with the 3xTF32 split
//   a = a_hi + a_lo, b = b_hi + b_lo (hi = top 19 bits), a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi, fp32 accumulate
// so results keep fp32-level accuracy (error ~2^-20 relative per product).  One CTA = tile row I, a run of up to
// PROP_RUN consecutive J tiles: the direct product Y[I rows] accumulates in registers over the run and is flushed
// once; the mirrored product Y[J rows] is flushed per tile.  Fragments are read straight from the staged tile:
//   direct  : A[m=i][k=j] = xs[i][j]          (k slots t, t+4  -> j = k0+t, k0+t+4)
//   mirrored: A[m=j][k=i] = xs[i][j]          (k slots t, t+4  -> i = k0+2t, k0+2t+1: conflict-free banks)
// ---------------------------------------------------------------------------------------------------------
constexpr int PROP_RUN = 4;
constexpr int BJ_LD_EXTRA = 8;    // bj row stride KC+8  -> banks 8t+g distinct for rows t, cols g
constexpr int BI_LD_EXTRA = 4;    // bi row stride KC+4  -> banks (2t)*(KC+4)+g = 8t+g (KC=32) distinct

template <int KC>
struct PropMmaSmem {
  float xs[TILE][XS_LD];
  float bj[TILE][KC + BJ_LD_EXTRA];
  float bi[TILE][KC + BI_LD_EXTRA];
  float rI[TILE], rJ[TILE], lseAI[TILE], lseAJ[TILE], lseFI[TILE], lseFJ[TILE], dlI[TILE], dlJ[TILE];
  float rowacc[TILE], colacc[TILE];
  double red[32];
};

__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(v) & 0xffffe000u;
  lo = __float_as_uint(v - __uint_as_float(hi));
}

__device__ __forceinline__ void mma_tf32(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// 256 threads: all 8 warps stage the tile (and do the fused element-wise terms); then warps 0-3 run the direct
// product (accumulating over the run) while warps 4-7 run the mirrored product of the same staged tile.
// MMAs are issued term-major over 8 independent accumulators so dependent MMAs are >= 8 instructions apart.
template <int KC, bool ELEM>
__global__ void __launch_bounds__(256, 2)
k_propagate_mma(const float* __restrict__ tiles, int64_t n, int tr0, const float* mu, int raw,
                const float* __restrict__ B, float* __restrict__ Y, mcgra_elem_args ea) {
  const int I = tr0 + (int)blockIdx.y;
  const int Jbeg = (int)blockIdx.x * PROP_RUN;
  if (Jbeg > I) return;
  const int Jend = min(I + 1, Jbeg + PROP_RUN);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PropMmaSmem<KC>& sm = *reinterpret_cast<PropMmaSmem<KC>*>(smem_raw);
  const ParamView pv = load_view(mu, raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int64_t i0 = (int64_t)I * TILE;
  constexpr int NB = KC / 8;
  const bool direct = warp < 4;
  const int wq = warp & 3;

  for (int e = tid; e < TILE * KC / 4; e += 256) {
    const int row = e / (KC / 4), c4 = e % (KC / 4);
    float4 vi = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i0 + row < n) vi = reinterpret_cast<const float4*>(B + (i0 + row) * KC)[c4];
    *reinterpret_cast<float4*>(&sm.bi[row][c4 * 4]) = vi;
  }
  if (ELEM && tid < TILE) {
    const int64_t gi = i0 + tid;
    sm.rI[tid] = gi < n ? ea.r[gi] : 0.f;
    if (ea.measure == MCGRA_M_KL) {
      sm.lseAI[tid] = gi < n ? ea.lseA[gi] : 0.f;
      sm.lseFI[tid] = gi < n ? ea.lseF[gi] : 0.f;
      sm.dlI[tid] = gi < n ? ea.dlse[gi] : 0.f;
    }
    sm.rowacc[tid] = 0.f;
  }
  float acc[2][NB][4];               // direct warps: running sum over the run; mirrored warps: per tile
#pragma unroll
  for (int mb = 0; mb < 2; ++mb)
#pragma unroll
    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[mb][nb][q] = 0.f;
  float v1 = 0.f, v6 = 0.f;

  for (int J = Jbeg; J < Jend; ++J) {
    const int64_t j0 = (int64_t)J * TILE;
    const int64_t tix = tri((int64_t)I) + J - tri((int64_t)tr0);
    const float4* src = reinterpret_cast<const float4*>(tiles + tix * TILE_ELEMS);
    __syncthreads();                       // previous tile's consumers are done with xs / bj
    for (int e = tid; e < TILE * KC / 4; e += 256) {
      const int row = e / (KC / 4), c4 = e % (KC / 4);
      float4 vj = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j0 + row < n) vj = reinterpret_cast<const float4*>(B + (j0 + row) * KC)[c4];
      *reinterpret_cast<float4*>(&sm.bj[row][c4 * 4]) = vj;
    }
    if (ELEM) {
      if (tid < TILE) {
        const int64_t gj = j0 + tid;
        sm.rJ[tid] = gj < n ? ea.r[gj] : 0.f;
        if (ea.measure == MCGRA_M_KL) {
          sm.lseAJ[tid] = gj < n ? ea.lseA[gj] : 0.f;
          sm.lseFJ[tid] = gj < n ? ea.lseF[gj] : 0.f;
          sm.dlJ[tid] = gj < n ? ea.dlse[gj] : 0.f;
        }
        sm.colacc[tid] = 0.f;
      }
      __syncthreads();
    }
    // ---- stage the tile (8 warps, one 512 B row each per step), fused element-wise terms ----
    float col_e[4] = {0.f, 0.f, 0.f, 0.f};
    const float4* fsrc = (ELEM && ea.Ftiles != nullptr) ? reinterpret_cast<const float4*>(ea.Ftiles + tix * TILE_ELEMS)
                                                        : nullptr;
    const bool interior = (J < I) && (i0 + TILE <= n);     // every entry valid
    float rj4[4] = {0.f, 0.f, 0.f, 0.f};
    if (ELEM) {
#pragma unroll
      for (int k = 0; k < 4; ++k) rj4[k] = sm.rJ[lane * 4 + k];
    }
#pragma unroll 4
    for (int it = 0; it < 16; ++it) {
      const int row = it * 8 + warp;
      const int idx = row * 32 + lane;
      const float4 raw4 = src[idx];
      float4 f4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ELEM && fsrc != nullptr) f4 = fsrc[idx];
      const int gi = (int)(i0 + row), gj = (int)(j0 + lane * 4);
      float xv[4] = {raw4.x, raw4.y, raw4.z, raw4.w};
      bool ok[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        ok[k] = interior || ((gj + k < gi) && (gi < n));
        xv[k] = ok[k] ? pv.adj(xv[k]) : 0.f;
      }
      *reinterpret_cast<float4*>(&sm.xs[row][lane * 4]) = make_float4(xv[0], xv[1], xv[2], xv[3]);
      if (ELEM) {
        const float fv[4] = {f4.x, f4.y, f4.z, f4.w};
        const float ri = sm.rI[row];
        float row_e = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (!ok[k]) continue;
          const float rj = rj4[k];
          const float ah = ri * xv[k] * rj;
          float esym = 0.f;   // e'_ij + e'_ji
          if (ea.measure == MCGRA_M_MSE) {
            const float df = ah - fv[k];
            v1 = fmaf(2.f * df, df, v1);
            esym = 4.f * ea.k1 * df;
          } else if (ea.measure == MCGRA_M_KL) {
            const float xij = __expf(fv[k] - sm.lseFI[row]);
            const float xji = __expf(fv[k] - sm.lseFJ[lane * 4 + k]);
            const float lij = ah - sm.lseAI[row];
            const float lji = ah - sm.lseAJ[lane * 4 + k];
            v1 += xij * ((fv[k] - ah) - sm.dlI[row]) + xji * ((fv[k] - ah) - sm.dlJ[lane * 4 + k]);
            esym = ea.k1 * ((__expf(lij) - xij) + (__expf(lji) - xji));
          }
          if (ea.k6 != 0.f) {
            const float q = fminf(fmaxf(ah, ENT_LO), ENT_HI);
            const float lg = __log2f(q);
            v6 = fmaf(2.f * q, lg, v6);
            if (ah >= ENT_LO && ah <= ENT_HI) esym = fmaf(2.f * ea.k6, lg + INV_LN2, esym);
          }
          const float tt = esym * xv[k];
          row_e = fmaf(tt, rj, row_e);
          col_e[k] = fmaf(tt, ri, col_e[k]);
        }
        row_e = warp_sum(row_e);
        if (lane == 0) sm.rowacc[row] += row_e;
      }
    }
    if (ELEM) {
#pragma unroll
      for (int k = 0; k < 4; ++k) atomicAdd(&sm.colacc[lane * 4 + k], col_e[k]);
    }
    __syncthreads();

    if (direct) {
      // ---- direct product: rows i in [32 wq, +32), accumulate over the run ----
#pragma unroll 2
      for (int ks = 0; ks < TILE / 8; ++ks) {
        const int k0 = ks * 8;
        uint32_t ahi[2][4], alo[2][4], bh[NB][2], bl[NB][2];
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
          const int m0 = wq * 32 + mb * 16;
          split_tf32(sm.xs[m0 + g][k0 + t], ahi[mb][0], alo[mb][0]);
          split_tf32(sm.xs[m0 + g + 8][k0 + t], ahi[mb][1], alo[mb][1]);
          split_tf32(sm.xs[m0 + g][k0 + t + 4], ahi[mb][2], alo[mb][2]);
          split_tf32(sm.xs[m0 + g + 8][k0 + t + 4], ahi[mb][3], alo[mb][3]);
        }
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
          split_tf32(sm.bj[k0 + t][nb * 8 + g], bh[nb][0], bl[nb][0]);
          split_tf32(sm.bj[k0 + t + 4][nb * 8 + g], bh[nb][1], bl[nb][1]);
        }
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
          for (int mb = 0; mb < 2; ++mb) mma_tf32(acc[mb][nb], alo[mb], bh[nb][0], bh[nb][1]);
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
          for (int mb = 0; mb < 2; ++mb) mma_tf32(acc[mb][nb], ahi[mb], bl[nb][0], bl[nb][1]);
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
          for (int mb = 0; mb < 2; ++mb) mma_tf32(acc[mb][nb], ahi[mb], bh[nb][0], bh[nb][1]);
      }
    } else {
      // ---- mirrored product: rows j in [32 wq, +32), flushed per tile ----
#pragma unroll 2
      for (int ks = 0; ks < TILE / 8; ++ks) {
        const int k0 = ks * 8;
        uint32_t ahi[2][4], alo[2][4], bh[NB][2], bl[NB][2];
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
          const int m0 = wq * 32 + mb * 16;
          split_tf32(sm.xs[k0 + 2 * t][m0 + g], ahi[mb][0], alo[mb][0]);
          split_tf32(sm.xs[k0 + 2 * t][m0 + g + 8], ahi[mb][1], alo[mb][1]);
          split_tf32(sm.xs[k0 + 2 * t + 1][m0 + g], ahi[mb][2], alo[mb][2]);
          split_tf32(sm.xs[k0 + 2 * t + 1][m0 + g + 8], ahi[mb][3], alo[mb][3]);
        }
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
          split_tf32(sm.bi[k0 + 2 * t][nb * 8 + g], bh[nb][0], bl[nb][0]);
          split_tf32(sm.bi[k0 + 2 * t + 1][nb * 8 + g], bh[nb][1], bl[nb][1]);
        }
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
          for (int mb = 0; mb < 2; ++mb) mma_tf32(acc[mb][nb], alo[mb], bh[nb][0], bh[nb][1]);
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
          for (int mb = 0; mb < 2; ++mb) mma_tf32(acc[mb][nb], ahi[mb], bl[nb][0], bl[nb][1]);
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
          for (int mb = 0; mb < 2; ++mb) mma_tf32(acc[mb][nb], ahi[mb], bh[nb][0], bh[nb][1]);
      }
#pragma unroll
      for (int mb = 0; mb < 2; ++mb) {
        const int64_t ra = j0 + wq * 32 + mb * 16 + g, rb = ra + 8;
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
          if (ra < n) atomicAdd(reinterpret_cast<float2*>(Y + ra * KC + nb * 8 + 2 * t),
                                make_float2(acc[mb][nb][0], acc[mb][nb][1]));
          if (rb < n) atomicAdd(reinterpret_cast<float2*>(Y + rb * KC + nb * 8 + 2 * t),
                                make_float2(acc[mb][nb][2], acc[mb][nb][3]));
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[mb][nb][q] = 0.f;
        }
      }
    }
    if (ELEM && tid < TILE) {
      const int64_t gj = j0 + tid;
      if (gj < n && sm.colacc[tid] != 0.f) atomicAdd(ea.eps_row + gj, sm.colacc[tid]);
    }
  }
  // ---- flush the direct product of the run ----
  if (direct) {
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) {
      const int64_t ra = i0 + wq * 32 + mb * 16 + g, rb = ra + 8;
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        if (ra < n) atomicAdd(reinterpret_cast<float2*>(Y + ra * KC + nb * 8 + 2 * t),
                              make_float2(acc[mb][nb][0], acc[mb][nb][1]));
        if (rb < n) atomicAdd(reinterpret_cast<float2*>(Y + rb * KC + nb * 8 + 2 * t),
                              make_float2(acc[mb][nb][2], acc[mb][nb][3]));
      }
    }
  }
  if (ELEM) {
    __syncthreads();
    if (tid < TILE) {
      const int64_t gi = i0 + tid;
      if (gi < n && sm.rowacc[tid] != 0.f) atomicAdd(ea.eps_row + gi, sm.rowacc[tid]);
    }
    if (ea.measure != MCGRA_M_NONE) block_atomic_add_d((double)v1 * (double)ea.k1, ea.acc + MCGRA_ACC_C1, sm.red);
    if (ea.k6 != 0.f) block_atomic_add_d((double)v6 * (double)ea.k6, ea.acc + MCGRA_ACC_C6, sm.red);
  }
}

template <int KC, bool ELEM>
int launch_prop_mma(const float* tiles, int64_t n, int tr0, int tr1, const float* mu, int raw, const float* B, float* Y,
                    const mcgra_elem_args* elem, cudaStream_t st) {
  const size_t smem = sizeof(PropMmaSmem<KC>);
  cudaError_t e = cudaFuncSetAttribute(k_propagate_mma<KC, ELEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  mcgra_elem_args ea = {};
  if (ELEM) ea = *elem;
  if (tr1 - tr0 > 65535) return -3;
  dim3 grid((unsigned)((tr1 + PROP_RUN - 1) / PROP_RUN), (unsigned)(tr1 - tr0));
  k_propagate_mma<KC, ELEM><<<grid, 256, smem, st>>>(tiles, n, tr0, mu, raw, B, Y, ea);
  MCGRA_LAUNCH_CHECK();
  return 0;
}

int g_prop_engine = 1;    // 0 = fp32 FFMA (v1), 1 = mma.sync 3xTF32 (v2)

template <int KC, bool ELEM>
int launch_prop(const float* tiles, int64_t n, int64_t t0, int64_t nt, const float* mu, int raw, const float* B,
                float* Y, const mcgra_elem_args* elem, cudaStream_t st) {
  const size_t smem = sizeof(PropSmem<KC>);
  cudaError_t e = cudaFuncSetAttribute(k_propagate<KC, ELEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  mcgra_elem_args ea = {};
  if (ELEM) ea = *elem;
  k_propagate<KC, ELEM><<<(unsigned)nt, 128, smem, st>>>(tiles, n, t0, mu, raw, B, Y, ea);
  MCGRA_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" {

int mcgra_set_fold_engine_(int value);
int mcgra_set_pairs_engine_(int value);
int mcgra_set_engine(int which, int value) {
  if (which == 0) { g_prop_engine = value; return 0; }
  if (which == 1) return mcgra_set_fold_engine_(value);
  if (which == 2) return mcgra_set_pairs_engine_(value);
  return -1;
}

int mcgra_degree(const float* tiles, int64_t n, int tr0, int tr1, const float* mu, int raw, float* d, void* stream) {
  const int64_t nt = tri(tr1) - tri(tr0);
  if (nt <= 0) return 0;
  k_degree<<<(unsigned)nt, 128, 0, (cudaStream_t)stream>>>(tiles, n, tri(tr0), mu, raw, d);
  MCGRA_LAUNCH_CHECK();
  return 0;
}

int mcgra_row_sumexp(const float* tiles, int64_t n, int tr0, int tr1, const float* mu, int raw, const float* r,
                     float* sumexp, void* stream) {
  const int64_t nt = tri(tr1) - tri(tr0);
  if (nt <= 0) return 0;
  k_row_sumexp<<<(unsigned)nt, 128, 0, (cudaStream_t)stream>>>(tiles, n, tri(tr0), mu, raw, r, sumexp);
  MCGRA_LAUNCH_CHECK();
  return 0;
}

int mcgra_propagate(const float* tiles, int64_t n, int tr0, int tr1, const float* mu, int raw, const float* B, int K,
                    float* Y, const mcgra_elem_args* elem, void* stream) {
  const int64_t nt = tri(tr1) - tri(tr0);
  if (nt <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t t0 = tri(tr0);
  if (g_prop_engine == 1) {
    if (K == 32)
      return elem ? launch_prop_mma<32, true>(tiles, n, tr0, tr1, mu, raw, B, Y, elem, st)
                  : launch_prop_mma<32, false>(tiles, n, tr0, tr1, mu, raw, B, Y, nullptr, st);
    if (K == 16)
      return elem ? launch_prop_mma<16, true>(tiles, n, tr0, tr1, mu, raw, B, Y, elem, st)
                  : launch_prop_mma<16, false>(tiles, n, tr0, tr1, mu, raw, B, Y, nullptr, st);
    return -1;
  }
  if (K == 32) {
    return elem ? launch_prop<32, true>(tiles, n, t0, nt, mu, raw, B, Y, elem, st)
                : launch_prop<32, false>(tiles, n, t0, nt, mu, raw, B, Y, nullptr, st);
  }
  if (K == 16) {
    return elem ? launch_prop<16, true>(tiles, n, t0, nt, mu, raw, B, Y, elem, st)
                : launch_prop<16, false>(tiles, n, t0, nt, mu, raw, B, Y, nullptr, st);
  }
  return -1;
}

}  // extern "C"
