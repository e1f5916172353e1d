// Exact-fp32 CUDA-core kernels for the "pairwise softmax" family (a1 global
// matching, a2 flow-propagation attention): S = X Y^T * scale between two token
// sets, row softmax, expectation of a 2-vector per column.
//
//   pair_fwd_simt : flash-style forward (online softmax, S never stored unless
//                   the caller asks for it = the `corr` output of matching.py:18-20)
//   pair_bwd_simt : backward for one operand: recomputes S tile by tile, forms
//                   the gradient tile W (see below) and accumulates dX = W Y * scale.
//
// These are the exact-arithmetic path: used for every backward pass and for
// forward shapes the tcgen05 kernel (match_tc.cu) does not cover.  All reductions
// are deterministic (no atomics).
//
// Gradient tile (rows r of X, columns c of Y), all terms optional:
//   W[r,c] = exp(S[r,c] - L1[r]) * (u[r].t[c]  - u0[r])      softmax over c (row direction)
//          + exp(S[r,c] - L2[c]) * (w[c].t2[r] - w0[c])      softmax over r (column direction)
//          + E[r*er + c*ec]                                  direct gradient on S (dcorr)
// with u,t,w,t2 2-vectors.  a1 backward (reference autograd of matching.py:16-39)
// and a2 backward (transformer.py:528-531, value detached) are both instances.
#include "common.cuh"
#include "../../include/emip_b200.h"
#include "pair_common.cuh"
#include "simt_tiles.cuh"

namespace {
using namespace simt;

// ------------------------------------------------------------------ forward
__global__ void __launch_bounds__(NT)
pair_fwd_simt_kernel(PairFwdArgs p) {
  extern __shared__ __align__(16) float smem[];
  float* Xs = smem;
  float* Ys = Xs + BM * LDX;
  float* Vs = Ys + BN * LDX;      // [2][BN]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int nrb = (p.nq + BM - 1) / BM;
  const int b = blockIdx.x / nrb, rb = blockIdx.x % nrb;
  const int row0 = rb * BM;
  const int bx = b, by = (b + p.y_shift) % p.nb;
  const float* X = p.x + (size_t)bx * p.nq * KC;
  const float* Y = p.y + (size_t)by * p.nk * KC;
  const float* V = p.v + (size_t)b * p.v_stride_b;

  load_tile(Xs, X, p.x_layout, row0, p.nq, p.nq);

  float m_run[4], l_run[4], sx[4], sy[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) { m_run[a] = -INFINITY; l_run[a] = 0.f; sx[a] = 0.f; sy[a] = 0.f; }

  for (int col0 = 0; col0 < p.nk; col0 += BN) {
    __syncthreads();   // previous tile fully consumed (also orders the X load on the first trip)
    load_tile(Ys, Y, p.y_layout, col0, p.nk, p.nk);
    if (tid < 2 * BN) {
      int ch = tid / BN, c = tid % BN;
      Vs[tid] = (col0 + c < p.nk) ? __ldg(V + (size_t)ch * p.nk + col0 + c) : 0.f;
    }
    __syncthreads();
    float acc[4][4];
    s_tile(Xs, Ys, tx, ty, acc);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      float tmax = -INFINITY;
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        int col = col0 + tx + 16 * bb;
        // matching.py:16 / transformer.py:528 divide by sqrt(C)
        float s = __fdiv_rn(acc[a][bb], p.sqrt_c);
        if (p.s_out != nullptr) {
          int row = row0 + ty + 16 * a;
          if (row < p.nq && col < p.nk) p.s_out[((size_t)b * p.nq + row) * p.nk + col] = s;
        }
        s = (col < p.nk) ? s : -INFINITY;
        acc[a][bb] = s;
        tmax = fmaxf(tmax, s);
      }
      tmax = group16_max(tmax);
      float m_new = fmaxf(m_run[a], tmax);     // finite: every tile has >= 1 valid column
      float corr = __expf(m_run[a] - m_new);   // exp(-inf) = 0 on the first tile
      float ls = 0.f, lx = 0.f, ly = 0.f;
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        float pr = expf(acc[a][bb] - m_new);
        ls += pr;
        lx = fmaf(pr, Vs[tx + 16 * bb], lx);
        ly = fmaf(pr, Vs[BN + tx + 16 * bb], ly);
      }
      l_run[a] = l_run[a] * corr + ls;
      sx[a] = sx[a] * corr + lx;
      sy[a] = sy[a] * corr + ly;
      m_run[a] = m_new;
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    float l = group16_sum(l_run[a]);
    float ex = group16_sum(sx[a]) / l;
    float ey = group16_sum(sy[a]) / l;
    int row = row0 + ty + 16 * a;
    if (tx == 0 && row < p.nq) {
      if (p.sub != nullptr) { ex -= __ldg(p.sub + row); ey -= __ldg(p.sub + p.nq + row); }
      p.out[((size_t)b * 2 + 0) * p.nq + row] = ex;
      p.out[((size_t)b * 2 + 1) * p.nq + row] = ey;
      if (p.lse != nullptr) p.lse[(size_t)b * p.nq + row] = m_run[a] + logf(l);
    }
  }
}

// ------------------------------------------------------------------ backward
__global__ void __launch_bounds__(NT)
pair_bwd_simt_kernel(PairBwdArgs p) {
  extern __shared__ __align__(16) float smem[];
  float* Xs = smem;                      // [BM][LDX]   (re-used to stage the transposed output)
  float* Ys = Xs + BM * LDX;             // [BN][LDX]
  float* Ws = Ys + BN * LDX;             // [BM][LDW]
  float* Cs = Ws + BM * LDW;             // [6][BN] per-column terms: tcx,tcy,L2,wx,wy,w0
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int nrb = (p.nr + BM - 1) / BM;
  const int b = blockIdx.x / nrb, rb = blockIdx.x % nrb;
  const int row0 = rb * BM;
  const int bx = (b + p.x_shift) % p.nb, by = (b + p.y_shift) % p.nb;
  const float* X = p.x + (size_t)bx * p.nr * KC;
  const float* Y = p.y + (size_t)by * p.nc * KC;
  const bool term1 = p.l1 != nullptr, term2 = p.l2 != nullptr, extra = p.e != nullptr;

  load_tile(Xs, X, p.x_layout, row0, p.nr, p.nr);

  // per-row terms for this thread's 4 rows
  float rL1[4], rux[4], ruy[4], ru0[4], rtx[4], rty[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    int row = row0 + ty + 16 * a;
    bool ok = row < p.nr;
    rL1[a] = rux[a] = ruy[a] = ru0[a] = rtx[a] = rty[a] = 0.f;
    if (ok && term1) {
      rL1[a] = __ldg(p.l1 + (size_t)b * p.nr + row);
      rux[a] = __ldg(p.u + ((size_t)b * 2 + 0) * p.nr + row);
      ruy[a] = __ldg(p.u + ((size_t)b * 2 + 1) * p.nr + row);
      ru0[a] = __ldg(p.u0 + (size_t)b * p.nr + row);
    }
    if (ok && term2) {
      rtx[a] = __ldg(p.t2 + (size_t)b * p.t2_stride_b + row);
      rty[a] = __ldg(p.t2 + (size_t)b * p.t2_stride_b + p.nr + row);
    }
  }

  float dacc[4][8];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int d = 0; d < 8; ++d) dacc[a][d] = 0.f;

  for (int col0 = 0; col0 < p.nc; col0 += BN) {
    __syncthreads();
    load_tile(Ys, Y, p.y_layout, col0, p.nc, p.nc);
    for (int i = tid; i < 6 * BN; i += NT) {
      int k = i / BN, c = i % BN, col = col0 + c;
      float v = 0.f;
      if (col < p.nc) {
        if (k < 2) { if (term1) v = __ldg(p.t + (size_t)b * p.t_stride_b + (size_t)k * p.nc + col); }
        else if (term2) {
          if (k == 2) v = __ldg(p.l2 + (size_t)b * p.nc + col);
          else if (k == 3) v = __ldg(p.w + ((size_t)b * 2 + 0) * p.nc + col);
          else if (k == 4) v = __ldg(p.w + ((size_t)b * 2 + 1) * p.nc + col);
          else v = __ldg(p.w0 + (size_t)b * p.nc + col);
        }
      }
      Cs[i] = v;
    }
    if (extra) {
      // stage E tile into Ws[row][col], coalesced along whichever index is contiguous in memory
      const float* E = p.e + (size_t)b * p.e_stride_b;
      if (p.e_stride_r == 1) {
        for (int i = tid; i < BM * BN; i += NT) {
          int c = i / BM, r = i % BM;
          int row = row0 + r, col = col0 + c;
          Ws[r * LDW + c] = (row < p.nr && col < p.nc) ? __ldg(E + (size_t)col * p.e_stride_c + row) : 0.f;
        }
      } else {
        for (int i = tid; i < BM * BN; i += NT) {
          int r = i / BN, c = i % BN;
          int row = row0 + r, col = col0 + c;
          Ws[r * LDW + c] = (row < p.nr && col < p.nc)
                                ? __ldg(E + (size_t)row * p.e_stride_r + (size_t)col * p.e_stride_c) : 0.f;
        }
      }
    }
    __syncthreads();
    float acc[4][4];
    s_tile(Xs, Ys, tx, ty, acc);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      int row = row0 + ty + 16 * a;
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        int c = tx + 16 * bb, col = col0 + c;
        float s = __fdiv_rn(acc[a][bb], p.sqrt_c);
        float w = 0.f;
        if (term1) w += expf(s - rL1[a]) * (rux[a] * Cs[c] + ruy[a] * Cs[BN + c] - ru0[a]);
        if (term2) w += expf(s - Cs[2 * BN + c]) * (Cs[3 * BN + c] * rtx[a] + Cs[4 * BN + c] * rty[a] - Cs[5 * BN + c]);
        if (extra) w += Ws[(ty + 16 * a) * LDW + c];
        if (row >= p.nr || col >= p.nc) w = 0.f;
        Ws[(ty + 16 * a) * LDW + c] = w;      // same thread staged/reads/writes this element
      }
    }
    __syncthreads();
    // dX[r][d] += sum_c W[r][c] * Y[c][d];  thread owns rows ty+16a, channels (tx+16e)*4..+3
#pragma unroll 2
    for (int c4 = 0; c4 < BN / 4; ++c4) {
      float4 wa[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) wa[a] = *reinterpret_cast<const float4*>(Ws + (ty + 16 * a) * LDW + c4 * 4);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float4 y0 = *reinterpret_cast<const float4*>(Ys + (c4 * 4 + k) * LDX + tx * 4);
        float4 y1 = *reinterpret_cast<const float4*>(Ys + (c4 * 4 + k) * LDX + (tx + 16) * 4);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          float wv = k == 0 ? wa[a].x : k == 1 ? wa[a].y : k == 2 ? wa[a].z : wa[a].w;
          dacc[a][0] = fmaf(wv, y0.x, dacc[a][0]);
          dacc[a][1] = fmaf(wv, y0.y, dacc[a][1]);
          dacc[a][2] = fmaf(wv, y0.z, dacc[a][2]);
          dacc[a][3] = fmaf(wv, y0.w, dacc[a][3]);
          dacc[a][4] = fmaf(wv, y1.x, dacc[a][4]);
          dacc[a][5] = fmaf(wv, y1.y, dacc[a][5]);
          dacc[a][6] = fmaf(wv, y1.z, dacc[a][6]);
          dacc[a][7] = fmaf(wv, y1.w, dacc[a][7]);
        }
      }
    }
  }
  // dS -> d(X): chain through the division by sqrt(C)
  const float inv = 1.0f / p.sqrt_c;
  float* DX = p.dx + (size_t)b * p.nr * KC;
  if (p.dx_layout == 0) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      int row = row0 + ty + 16 * a;
      if (row < p.nr) {
        float4 o0 = make_float4(dacc[a][0] * inv, dacc[a][1] * inv, dacc[a][2] * inv, dacc[a][3] * inv);
        float4 o1 = make_float4(dacc[a][4] * inv, dacc[a][5] * inv, dacc[a][6] * inv, dacc[a][7] * inv);
        float4* dst = reinterpret_cast<float4*>(DX + (size_t)row * KC);
        if (p.accumulate) {
          float4 q0 = dst[tx], q1 = dst[tx + 16];
          o0.x += q0.x; o0.y += q0.y; o0.z += q0.z; o0.w += q0.w;
          o1.x += q1.x; o1.y += q1.y; o1.z += q1.z; o1.w += q1.w;
        }
        dst[tx] = o0;
        dst[tx + 16] = o1;
      }
    }
  } else {
    __syncthreads();   // all warps done with Xs (S tiles) before it is overwritten
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      float* dst = Xs + (ty + 16 * a) * LDX;
      *reinterpret_cast<float4*>(dst + tx * 4) =
          make_float4(dacc[a][0] * inv, dacc[a][1] * inv, dacc[a][2] * inv, dacc[a][3] * inv);
      *reinterpret_cast<float4*>(dst + (tx + 16) * 4) =
          make_float4(dacc[a][4] * inv, dacc[a][5] * inv, dacc[a][6] * inv, dacc[a][7] * inv);
    }
    __syncthreads();
    for (int i = tid; i < BM * KC; i += NT) {
      int c = i / BM, r = i % BM;
      int row = row0 + r;
      if (row < p.nr) {
        float* dst = DX + (size_t)c * p.nr + row;
        float v = Xs[r * LDX + c];
        *dst = p.accumulate ? (*dst + v) : v;
      }
    }
  }
}

// u0[b,i] = u[b,0,i]*(o[b,0,i]+g[0,i]) + u[b,1,i]*(o[b,1,i]+g[1,i])   (g optional)
__global__ void rowdot2_kernel(const float* __restrict__ u, const float* __restrict__ o, const float* __restrict__ g,
                               float* __restrict__ out, int nb, int n) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)nb * n) return;
  int b = (int)(idx / n), i = (int)(idx % n);
  float ox = o[((size_t)b * 2 + 0) * n + i], oy = o[((size_t)b * 2 + 1) * n + i];
  if (g != nullptr) { ox += g[i]; oy += g[n + i]; }
  out[idx] = u[((size_t)b * 2 + 0) * n + i] * ox + u[((size_t)b * 2 + 1) * n + i] * oy;
}

__global__ void coords_grid_kernel(float* __restrict__ g, int h, int w) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= h * w) return;
  g[j] = (float)(j % w);           // geometry.py:8 stacks [x, y]
  g[h * w + j] = (float)(j / w);
}

}  // namespace

size_t pair_fwd_simt_smem() { return sizeof(float) * (BM * LDX + BN * LDX + 2 * BN); }
size_t pair_bwd_simt_smem() { return sizeof(float) * (BM * LDX + BN * LDX + BM * LDW + 6 * BN); }

int pair_fwd_simt(const PairFwdArgs& a, cudaStream_t st) {
  if (a.nb == 0 || a.nq == 0) return EMIP_OK;
  if (int rc__ = emip_func_max_smem((const void*)(pair_fwd_simt_kernel), (int)pair_fwd_simt_smem())) return rc__;
  int nrb = (a.nq + BM - 1) / BM;
  pair_fwd_simt_kernel<<<a.nb * nrb, NT, pair_fwd_simt_smem(), st>>>(a);
  EMIP_CHECK_LAUNCH("pair_fwd_simt");
  return EMIP_OK;
}

int pair_bwd_simt(const PairBwdArgs& a, cudaStream_t st) {
  if (a.nb == 0 || a.nr == 0) return EMIP_OK;
  if (int rc__ = emip_func_max_smem((const void*)(pair_bwd_simt_kernel), (int)pair_bwd_simt_smem())) return rc__;
  int nrb = (a.nr + BM - 1) / BM;
  pair_bwd_simt_kernel<<<a.nb * nrb, NT, pair_bwd_simt_smem(), st>>>(a);
  EMIP_CHECK_LAUNCH("pair_bwd_simt");
  return EMIP_OK;
}

int launch_rowdot2(const float* u, const float* o, const float* g, float* out, int nb, int n, cudaStream_t st) {
  long long tot = (long long)nb * n;
  if (tot == 0) return EMIP_OK;
  rowdot2_kernel<<<(int)((tot + 255) / 256), 256, 0, st>>>(u, o, g, out, nb, n);
  EMIP_CHECK_LAUNCH("rowdot2");
  return EMIP_OK;
}

int launch_coords_grid(float* g, int h, int w, cudaStream_t st) {
  coords_grid_kernel<<<(h * w + 255) / 256, 256, 0, st>>>(g, h, w);
  EMIP_CHECK_LAUNCH("coords_grid");
  return EMIP_OK;
}
