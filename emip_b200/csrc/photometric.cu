// f3 (second half): the photometric term of the unsupervised flow loss, fused.
//
// Replaces reference loss/loss_flow.py:35-49 `unFlowLoss.loss_photomatric(im1_scaled, im1_recons, occu_mask1)` with
// loss/loss_blocks.py:46-65 `SSIM` (3x3 mean filters, no padding) for the configuration the reference trains
// (w_ternary = 0):
//   loss = ( w_l1 * mean(|im - rec| * m) + w_ssim * mean(clamp((1 - SSIM(rec*m, im*m)) / 2, 0, 1)) ) / mean(m)
// The reference runs ~25 elementwise / pooling kernels forward and twice that backward over [B,3,H,W] tensors; here the
// forward is one pass over the three inputs plus a tiny deterministic reduction, and the backward (to `rec`, the only
// input that carries a gradient: loss_flow.py:90-91) is one pass that recomputes the window statistics:
//   SSIM_p = n1 n2 / (d1 d2),  n1 = 2 mx my + C1, n2 = 2 sxy + C2, d1 = mx^2 + my^2 + C1, d2 = sx + sy + C2
//   dSSIM_p / dx_q = (2/9) (A_p + B_p y_q + C_p x_q)   for q in the 3x3 window of p, with
//   A_p = my (n2 - n1) / (d1 d2) - SSIM_p mx (1/d1 - 1/d2),  B_p = n1 / (d1 d2),  C_p = -SSIM_p / d2
// so every pixel q gathers the three coefficient sums of its (up to nine) windows -- no atomics, no scratch tensors.
#include "common.cuh"
#include "../../include/emip_b200.h"

namespace {
constexpr int TX = 32, TY = 32;           // pixel tile per CTA: 256 threads = 32 columns x 8 row groups of 4 consecutive rows
constexpr int NT = 256, RPT = 4;
constexpr float SSIM_C1 = 0.01f * 0.01f, SSIM_C2 = 0.03f * 0.03f;

struct PhotoArgs {
  const float* im; const float* rec; const float* mask;
  int B, C, H, W;
  float w_l1, w_ssim;
};

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  if (warp == 0) {
    t = lane < (int)(blockDim.x >> 5) ? red[lane] : 0.f;
    t = warp_sum(t);
  }
  return t;      // valid in warp 0
}

// horizontal 3-sums of one tile row at column c (c = left column of the window): (x, y, xx, yy, xy)
struct Row5 { float x, y, xx, yy, xy; };
__device__ __forceinline__ Row5 row5(const float* __restrict__ sx, const float* __restrict__ sy, int idx) {
  const float x0 = sx[idx], x1 = sx[idx + 1], x2 = sx[idx + 2], y0 = sy[idx], y1 = sy[idx + 1], y2 = sy[idx + 2];
  Row5 r;
  r.x = x0 + x1 + x2;
  r.y = y0 + y1 + y2;
  r.xx = fmaf(x0, x0, fmaf(x1, x1, x2 * x2));
  r.yy = fmaf(y0, y0, fmaf(y1, y1, y2 * y2));
  r.xy = fmaf(x0, y0, fmaf(x1, y1, x2 * y2));
  return r;
}
struct Stats { float mx, my, n1, n2, d1, d2, ssim; };
__device__ __forceinline__ Stats stats_of(const Row5& a, const Row5& b, const Row5& c) {
  const float k = 1.0f / 9.0f;
  Stats s;
  s.mx = (a.x + b.x + c.x) * k;
  s.my = (a.y + b.y + c.y) * k;
  const float sig_x = (a.xx + b.xx + c.xx) * k - s.mx * s.mx, sig_y = (a.yy + b.yy + c.yy) * k - s.my * s.my,
              sig_xy = (a.xy + b.xy + c.xy) * k - s.mx * s.my;
  s.n1 = 2.f * s.mx * s.my + SSIM_C1;
  s.n2 = 2.f * sig_xy + SSIM_C2;
  s.d1 = s.mx * s.mx + s.my * s.my + SSIM_C1;
  s.d2 = sig_x + sig_y + SSIM_C2;
  s.ssim = (s.n1 * s.n2) / (s.d1 * s.d2);
  return s;
}

// masked tiles x = rec * m, y = im * m of one channel with a HALO-pixel border (zero outside the image)
template <int HALO>
__device__ __forceinline__ void load_tiles(const float* __restrict__ im, const float* __restrict__ rc, const float* __restrict__ sm,
                                           float* __restrict__ sx, float* __restrict__ sy, int x0, int y0, int H, int W) {
  constexpr int SW = TX + 2 * HALO, SH = TY + 2 * HALO;
  for (int i = threadIdx.x; i < SH * SW; i += NT) {
    const int yy = y0 - HALO + i / SW, xx = x0 - HALO + i % SW;
    float a = 0.f, r = 0.f;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
      a = __ldg(im + (size_t)yy * W + xx);
      r = __ldg(rc + (size_t)yy * W + xx);
    }
    sx[i] = r * sm[i];
    sy[i] = a * sm[i];
  }
}
template <int HALO>
__device__ __forceinline__ void load_mask(const float* __restrict__ mask, float* __restrict__ sm, int x0, int y0, int H, int W) {
  constexpr int SW = TX + 2 * HALO, SH = TY + 2 * HALO;
  for (int i = threadIdx.x; i < SH * SW; i += NT) {
    const int yy = y0 - HALO + i / SW, xx = x0 - HALO + i % SW;
    sm[i] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(mask + (size_t)yy * W + xx) : 0.f;
  }
}

// partial[block] = (sum |im - rec| m, sum dist, sum m).  Every thread owns 4 consecutive rows of one column, so the
// six horizontal row sums it needs are shared by its four 3x3 windows.
__global__ void __launch_bounds__(NT)
photo_fwd_kernel(PhotoArgs p, float* __restrict__ partial) {
  constexpr int SW = TX + 2, SH = TY + 2;
  __shared__ float sx[SH * SW], sy[SH * SW], sm[SH * SW], red[8];
  const int b = blockIdx.z, x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
  const int tx = threadIdx.x % TX, tr = (threadIdx.x / TX) * RPT;      // first of my 4 rows (tile coordinates)
  const int gx = x0 + tx;
  const size_t plane = (size_t)p.H * p.W;
  load_mask<1>(p.mask + (size_t)b * plane, sm, x0, y0, p.H, p.W);
  float l1 = 0.f, ds = 0.f;
  for (int c = 0; c < p.C; ++c) {
    __syncthreads();                        // mask tile visible / previous channel consumed
    load_tiles<1>(p.im + ((size_t)b * p.C + c) * plane, p.rec + ((size_t)b * p.C + c) * plane, sm, sx, sy, x0, y0, p.H, p.W);
    __syncthreads();
    if (gx < p.W) {
      Row5 r0 = row5(sx, sy, tr * SW + tx), r1 = row5(sx, sy, (tr + 1) * SW + tx);
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        const Row5 r2 = row5(sx, sy, (tr + j + 2) * SW + tx);
        const int gy = y0 + tr + j;
        if (gy < p.H) {
          const int ci = (tr + j + 1) * SW + tx + 1;
          l1 += fabsf(sy[ci] - sx[ci]);                      // |im m - rec m| = |im - rec| m  (m >= 0)
          if (gx >= 1 && gx < p.W - 1 && gy >= 1 && gy < p.H - 1) {
            const Stats s = stats_of(r0, r1, r2);
            ds += fminf(fmaxf((1.f - s.ssim) * 0.5f, 0.f), 1.f);
          }
        }
        r0 = r1; r1 = r2;
      }
    }
  }
  float msum = 0.f;
  if (gx < p.W)
#pragma unroll
    for (int j = 0; j < RPT; ++j)
      if (y0 + tr + j < p.H) msum += sm[(tr + j + 1) * SW + tx + 1];
  const float t0 = block_sum(l1, red), t1 = block_sum(ds, red), t2 = block_sum(msum, red);
  if (threadIdx.x == 0) {
    const size_t blk = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    partial[3 * blk] = t0; partial[3 * blk + 1] = t1; partial[3 * blk + 2] = t2;
  }
}

// sums[0..2] = fixed-order totals, sums[3] = loss
__global__ void __launch_bounds__(1024)
photo_final_kernel(const float* __restrict__ partial, size_t n_blocks, float* __restrict__ sums, float* __restrict__ loss,
                   float w_l1, float w_ssim, double n1, double n2, double nm) {
  __shared__ double red[3][32];
  double a[3] = {0.0, 0.0, 0.0};
  for (size_t i = threadIdx.x; i < n_blocks; i += blockDim.x) {
    a[0] += partial[3 * i]; a[1] += partial[3 * i + 1]; a[2] += partial[3 * i + 2];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
    if (lane == 0) red[k][warp] = a[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[3] = {0.0, 0.0, 0.0};
    for (int w = 0; w < 32; ++w) { t[0] += red[0][w]; t[1] += red[1][w]; t[2] += red[2][w]; }
    sums[0] = (float)t[0]; sums[1] = (float)t[1]; sums[2] = (float)t[2];
    const double l = (w_l1 * t[0] / n1 + w_ssim * t[1] / n2) / (t[2] / nm);
    sums[3] = (float)l;
    if (loss != nullptr) *loss = (float)l;
  }
}

// drec = gl / mean(m) * ( w_l1 / n1 * d|im - rec| m / drec + w_ssim / n2 * d dist / drec )
// Phase 1: the (A, B, C) coefficients of every window centre in the tile + 1-pixel halo (34 x 34 centres; a thread
// owns 5 consecutive centres of one column and slides the row sums).  Phase 2: every pixel gathers the coefficient
// sums of its 3x3 neighbourhood, again through shared horizontal sums of its four consecutive rows.
__global__ void __launch_bounds__(NT)
photo_bwd_kernel(PhotoArgs p, const float* __restrict__ sums, const float* __restrict__ gloss, float* __restrict__ drec,
                 float inv_n1, float inv_n2, float nm) {
  constexpr int SW = TX + 4, SH = TY + 4;          // inputs with a 2-pixel halo
  constexpr int CW = TX + 2, CH = TY + 2;          // coefficients with a 1-pixel halo
  constexpr int CPT = 5;                           // centres per thread in phase 1: 7 row groups x 34 columns = 238 threads
  __shared__ float sx[SH * SW], sy[SH * SW], sm[SH * SW];
  __shared__ float ca[CH * CW], cb[CH * CW], cc[CH * CW];
  const int b = blockIdx.z, x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
  const int tx = threadIdx.x % TX, tr = (threadIdx.x / TX) * RPT;
  const int gx = x0 + tx;
  const size_t plane = (size_t)p.H * p.W;
  const float scale = __ldg(gloss) / (__ldg(sums + 2) / nm);
  const float k_l1 = scale * p.w_l1 * inv_n1;
  const float k_ss = scale * p.w_ssim * inv_n2 * (-0.5f) * (2.0f / 9.0f);
  load_mask<2>(p.mask + (size_t)b * plane, sm, x0, y0, p.H, p.W);
  const int ccol = threadIdx.x % CW, crow0 = (threadIdx.x / CW) * CPT;   // phase-1 assignment
  for (int c = 0; c < p.C; ++c) {
    __syncthreads();
    load_tiles<2>(p.im + ((size_t)b * p.C + c) * plane, p.rec + ((size_t)b * p.C + c) * plane, sm, sx, sy, x0, y0, p.H, p.W);
    __syncthreads();
    if (crow0 < CH) {
      // centre (crow, ccol) of the coefficient grid = pixel (y0 - 1 + crow, x0 - 1 + ccol) = input tile (crow + 1, ccol + 1)
      Row5 r0 = row5(sx, sy, crow0 * SW + ccol), r1 = row5(sx, sy, (crow0 + 1) * SW + ccol);
      const int px = x0 - 1 + ccol;
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const int crow = crow0 + j;
        if (crow < CH) {
          const Row5 r2 = row5(sx, sy, (crow + 2) * SW + ccol);
          const int py = y0 - 1 + crow;
          float A = 0.f, Bc = 0.f, Cc = 0.f;
          if (px >= 1 && px < p.W - 1 && py >= 1 && py < p.H - 1) {
            const Stats s = stats_of(r0, r1, r2);
            const float v = (1.f - s.ssim) * 0.5f;
            if (v >= 0.f && v <= 1.f) {                  // clamp of loss_blocks.py:64 active
              const float inv_dd = 1.f / (s.d1 * s.d2);
              A = s.my * (s.n2 - s.n1) * inv_dd - s.ssim * s.mx * (1.f / s.d1 - 1.f / s.d2);
              Bc = s.n1 * inv_dd;
              Cc = -s.ssim / s.d2;
            }
          }
          ca[crow * CW + ccol] = A; cb[crow * CW + ccol] = Bc; cc[crow * CW + ccol] = Cc;
          r0 = r1; r1 = r2;
        }
      }
    }
    __syncthreads();
    if (gx < p.W) {
      auto h3 = [&](const float* t, int row) { const int i = row * CW + tx; return t[i] + t[i + 1] + t[i + 2]; };
      float a0 = h3(ca, tr), a1 = h3(ca, tr + 1), b0 = h3(cb, tr), b1 = h3(cb, tr + 1), c0 = h3(cc, tr), c1 = h3(cc, tr + 1);
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        const float a2 = h3(ca, tr + j + 2), b2 = h3(cb, tr + j + 2), c2 = h3(cc, tr + j + 2);
        const int gy = y0 + tr + j;
        if (gy < p.H) {
          const int ci = (tr + j + 2) * SW + tx + 2;
          const float m = sm[ci], xq = sx[ci], yq = sy[ci];
          const float d = yq - xq;                           // (im - rec) m: same sign as im - rec where m > 0
          const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
          const float g = -k_l1 * sgn * m + k_ss * ((a0 + a1 + a2) + (b0 + b1 + b2) * yq + (c0 + c1 + c2) * xq) * m;
          drec[((size_t)b * p.C + c) * plane + (size_t)gy * p.W + gx] = g;
        }
        a0 = a1; a1 = a2; b0 = b1; b1 = b2; c0 = c1; c1 = c2;
      }
    }
  }
}
}  // namespace

extern "C" size_t emip_photometric_workspace(int B, int H, int W) {
  if (B < 0 || H <= 0 || W <= 0) return 0;
  const size_t blocks = (size_t)B * ((H + TY - 1) / TY) * ((W + TX - 1) / TX);
  return emip_align_up(blocks * 3 * sizeof(float), 256);
}

extern "C" int emip_photometric_fwd(const float* im, const float* rec, const float* mask, float* loss, float* sums,
                                    void* workspace, size_t ws_bytes, int B, int C, int H, int W, float w_l1, float w_ssim,
                                    void* stream) {
  EMIP_CHECK_ARG(im && rec && mask && sums && workspace, "photometric_fwd: null pointer");
  EMIP_CHECK_ARG(B > 0 && C > 0 && H >= 3 && W >= 3, "photometric_fwd: bad shape B=%d C=%d H=%d W=%d", B, C, H, W);
  if (ws_bytes < emip_photometric_workspace(B, H, W)) {
    emip_set_error("photometric_fwd: workspace too small");
    return EMIP_ENOMEM;
  }
  cudaStream_t st = (cudaStream_t)stream;
  PhotoArgs p = {im, rec, mask, B, C, H, W, w_l1, w_ssim};
  const dim3 grid((W + TX - 1) / TX, (H + TY - 1) / TY, B);
  EMIP_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "photometric_fwd: problem too large");
  float* partial = static_cast<float*>(workspace);
  photo_fwd_kernel<<<grid, NT, 0, st>>>(p, partial);
  EMIP_CHECK_LAUNCH("photometric_fwd");
  photo_final_kernel<<<1, 1024, 0, st>>>(partial, (size_t)grid.x * grid.y * grid.z, sums, loss, w_l1, w_ssim,
                                         (double)B * C * H * W, (double)B * C * (H - 2) * (W - 2), (double)B * H * W);
  EMIP_CHECK_LAUNCH("photometric_fwd (reduce)");
  return EMIP_OK;
}

extern "C" int emip_photometric_bwd(const float* im, const float* rec, const float* mask, const float* sums,
                                    const float* gloss, float* drec, int B, int C, int H, int W, float w_l1, float w_ssim,
                                    void* stream) {
  EMIP_CHECK_ARG(im && rec && mask && sums && gloss && drec, "photometric_bwd: null pointer");
  EMIP_CHECK_ARG(B > 0 && C > 0 && H >= 3 && W >= 3, "photometric_bwd: bad shape B=%d C=%d H=%d W=%d", B, C, H, W);
  cudaStream_t st = (cudaStream_t)stream;
  PhotoArgs p = {im, rec, mask, B, C, H, W, w_l1, w_ssim};
  const dim3 grid((W + TX - 1) / TX, (H + TY - 1) / TY, B);
  EMIP_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "photometric_bwd: problem too large");
  photo_bwd_kernel<<<grid, NT, 0, st>>>(p, sums, gloss, drec, (float)(1.0 / ((double)B * C * H * W)),
                                             (float)(1.0 / ((double)B * C * (H - 2) * (W - 2))), (float)((double)B * H * W));
  EMIP_CHECK_LAUNCH("photometric_bwd");
  return EMIP_OK;
}
