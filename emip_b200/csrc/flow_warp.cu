// K3: photometric-loss bilinear warp (flow_warp) forward and backward.
//
// Replaces reference loss/warp_utils.py:83-93 (mesh_grid :7-13, norm_grid :16-23,
// F.grid_sample bilinear / align_corners=True / border|zeros).  One fused pass:
// no base grid, no normalised grid tensor, no permute -- flow is read once
// (through explicit batch/channel strides, so the [B,4,H,W] channel slices of
// loss_flow.py:90-91 need no copy), the image taps come through the read-only
// path, the output is written once.  HBM-bound: algorithmic bytes per launch
// fwd = B*H*W*(2*C+2)*4, bwd-to-flow = B*H*W*(2*C+4)*4.
//
// The fp32 normalise -> un-normalise round trip of the reference
// (nx = 2u/(W-1)-1 ; ix = ((nx+1)/2)(W-1)) is reproduced in the same order.
#include "common.cuh"
#include "flow_warp_common.cuh"
#include "flow_warp_staged.cuh"
#include "../../include/emip_b200.h"

namespace {

struct Tap {
  int o00, o01, o10, o11;      // plane offsets of the four taps (clamped to be safe)
  float w00, w01, w10, w11;    // bilinear weights with out-of-range taps zeroed
  int valid;                   // bit k set when tap k (00,01,10,11) lies inside the image
};

struct Coord {
  float ix, iy;                // sample point in pixel units (after border clamp)
  float gmx, gmy;              // d(ix)/d(u) mask (0 where clamped), excluding the scale
};

template <bool BORDER>
__device__ __forceinline__ Coord sample_point(float u, float v, int H, int W) {
  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  // warp_utils.py:21-22
  float nx = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, u), wm1), 1.0f);
  float ny = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, v), hm1), 1.0f);
  // ATen grid_sampler_unnormalize, align_corners=True
  float ix = __fmul_rn(__fmul_rn(__fadd_rn(nx, 1.0f), 0.5f), wm1);
  float iy = __fmul_rn(__fmul_rn(__fadd_rn(ny, 1.0f), 0.5f), hm1);
  Coord c;
  c.gmx = 1.0f;
  c.gmy = 1.0f;
  if (BORDER) {
    // ATen clip_coordinates(_set_grad): gradient is zero on and outside the border
    if (!(ix > 0.0f)) { ix = 0.0f; c.gmx = 0.0f; }
    else if (ix >= wm1) { ix = wm1; c.gmx = 0.0f; }
    if (!(iy > 0.0f)) { iy = 0.0f; c.gmy = 0.0f; }
    else if (iy >= hm1) { iy = hm1; c.gmy = 0.0f; }
  }
  c.ix = ix;
  c.iy = iy;
  return c;
}

__device__ __forceinline__ Tap make_tap(const Coord& c, int H, int W, float& wx, float& wy) {
  float x0 = floorf(c.ix), y0 = floorf(c.iy);
  wx = c.ix - x0;
  wy = c.iy - y0;
  float ex = 1.0f - wx, ey = 1.0f - wy;
  float x1 = x0 + 1.0f, y1 = y0 + 1.0f;
  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  bool vx0 = (x0 >= 0.0f) && (x0 <= wm1), vx1 = (x1 >= 0.0f) && (x1 <= wm1);
  bool vy0 = (y0 >= 0.0f) && (y0 <= hm1), vy1 = (y1 >= 0.0f) && (y1 <= hm1);
  int xi0 = vx0 ? (int)x0 : 0, xi1 = vx1 ? (int)x1 : 0;
  int yi0 = vy0 ? (int)y0 : 0, yi1 = vy1 ? (int)y1 : 0;
  Tap t;
  t.o00 = yi0 * W + xi0;
  t.o01 = yi0 * W + xi1;
  t.o10 = yi1 * W + xi0;
  t.o11 = yi1 * W + xi1;
  t.w00 = (vx0 && vy0) ? ey * ex : 0.0f;
  t.w01 = (vx1 && vy0) ? ey * wx : 0.0f;
  t.w10 = (vx0 && vy1) ? wy * ex : 0.0f;
  t.w11 = (vx1 && vy1) ? wy * wx : 0.0f;
  t.valid = (int)(vx0 && vy0) | ((int)(vx1 && vy0) << 1) | ((int)(vx0 && vy1) << 2) | ((int)(vx1 && vy1) << 3);
  return t;
}

// Forward.  CTA = 256 threads = 32 x 8 lanes covering a 32 x 32 pixel tile of one sample: every thread owns
// ROWS = 4 pixels (rows 8 apart), all C channels.  Consecutive lanes = consecutive pixels, so for the piecewise-smooth
// flows the model produces the four tap requests of a warp touch 1-2 cache lines each (ncu r1a: a 4-consecutive-
// pixels-per-thread mapping made every tap request span ~30 sectors and held the kernel to 19 % of HBM peak).
// All 2*ROWS flow loads are issued first, then all 4*C*ROWS gathers, then the stores: ~50 independent loads in
// flight per thread keep HBM busy at 16 resident warps per SM.  Vertically adjacent rows of the tile share tap lines
// through L1 (taps come through the read-only path and are allowed to allocate in L1).
using k3::ROWS;
using k3::FastCoord;
using k3::fast_coord;
using k3::div_rn;
using k3::opaque;

template <bool BORDER, int CT>
__global__ void __launch_bounds__(256)
flow_warp_fwd_kernel(const float* __restrict__ x, const float* __restrict__ flow, float* __restrict__ out,
                     int C, int H, int W, long long fsb, long long fsc, int tiles_x, int tiles_y) {
  const long long plane = (long long)H * W;
  int t = blockIdx.x;
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y;
  const int b = t / tiles_y;
  const int px = tx * 32 + (threadIdx.x & 31);
  const int py0 = ty * (8 * ROWS) + (threadIdx.x >> 5);
  if (px >= W) return;
  const float* fb = flow + (long long)b * fsb + px;
  const float* xb = x + (long long)b * C * plane;
  float* ob = out + (long long)b * C * plane + px;
  float fx[ROWS], fy[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const int py = py0 + 8 * r;
    const bool ok = py < H;
    fx[r] = ok ? __ldcs(fb + (long long)py * W) : 0.f;            // streamed: read exactly once
    fy[r] = ok ? __ldcs(fb + fsc + (long long)py * W) : 0.f;
  }
  Tap tap[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const int py = min(py0 + 8 * r, H - 1);
    Coord c = sample_point<BORDER>((float)px + fx[r], (float)py + fy[r], H, W);
    float wx, wy;
    tap[r] = make_tap(c, H, W, wx, wy);
  }
  if constexpr (CT > 0) {
    float v[ROWS][CT > 0 ? CT : 1][4];
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
      for (int ch = 0; ch < CT; ++ch) {                           // every gather in flight before any use
        const float* xp = xb + ch * plane;
        v[r][ch][0] = __ldg(xp + tap[r].o00); v[r][ch][1] = __ldg(xp + tap[r].o01);
        v[r][ch][2] = __ldg(xp + tap[r].o10); v[r][ch][3] = __ldg(xp + tap[r].o11);
      }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const int py = py0 + 8 * r;
      if (py < H) {
#pragma unroll
        for (int ch = 0; ch < CT; ++ch)
          __stcs(ob + ch * plane + (long long)py * W, v[r][ch][0] * tap[r].w00 + v[r][ch][1] * tap[r].w01 +
                                                          v[r][ch][2] * tap[r].w10 + v[r][ch][3] * tap[r].w11);
      }
    }
  } else {
    for (int ch = 0; ch < C; ++ch) {
      const float* xp = xb + ch * plane;
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        const int py = py0 + 8 * r;
        float v00 = __ldg(xp + tap[r].o00), v01 = __ldg(xp + tap[r].o01), v10 = __ldg(xp + tap[r].o10),
              v11 = __ldg(xp + tap[r].o11);
        if (py < H)
          __stcs(ob + ch * plane + (long long)py * W, v00 * tap[r].w00 + v01 * tap[r].w01 + v10 * tap[r].w10 + v11 * tap[r].w11);
      }
    }
  }
}

// Backward: dflow per pixel (no atomics); optional dx by atomic scatter.  Same tiling as the forward.
template <bool BORDER, bool WITH_DX, int CT>
__global__ void __launch_bounds__(256)
flow_warp_bwd_kernel(const float* __restrict__ x, const float* __restrict__ flow, const float* __restrict__ dout,
                     float* __restrict__ dflow, float* __restrict__ dx,
                     int C, int H, int W, long long fsb, long long fsc, int tiles_x, int tiles_y) {
  const long long plane = (long long)H * W;
  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  int t = blockIdx.x;
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y;
  const int b = t / tiles_y;
  const int px = tx * 32 + (threadIdx.x & 31);
  const int py0 = ty * (8 * ROWS) + (threadIdx.x >> 5);
  if (px >= W) return;
  const float* fb = flow + (long long)b * fsb + px;
  const float* xb = x + (long long)b * C * plane;
  const float* gb = dout + (long long)b * C * plane + px;
  float fx[ROWS], fy[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const int py = py0 + 8 * r;
    const bool ok = py < H;
    fx[r] = ok ? __ldcs(fb + (long long)py * W) : 0.f;
    fy[r] = ok ? __ldcs(fb + fsc + (long long)py * W) : 0.f;
  }
  if constexpr (CT > 0 && !WITH_DX) {
    // the photometric-loss case (images need no gradient): every load of the 4 rows in flight before any use
    Tap tap[ROWS];
    Coord co[ROWS];
    float wxs[ROWS], wys[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const int py = min(py0 + 8 * r, H - 1);
      co[r] = sample_point<BORDER>((float)px + fx[r], (float)py + fy[r], H, W);
      tap[r] = make_tap(co[r], H, W, wxs[r], wys[r]);
    }
    float g[ROWS][CT > 0 ? CT : 1], v[ROWS][CT > 0 ? CT : 1][4];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const long long pix = (long long)min(py0 + 8 * r, H - 1) * W;
#pragma unroll
      for (int ch = 0; ch < CT; ++ch) {
        const float* xp = xb + ch * plane;
        g[r][ch] = __ldcs(gb + ch * plane + pix);
        v[r][ch][0] = __ldg(xp + tap[r].o00); v[r][ch][1] = __ldg(xp + tap[r].o01);
        v[r][ch][2] = __ldg(xp + tap[r].o10); v[r][ch][3] = __ldg(xp + tap[r].o11);
      }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const int py = py0 + 8 * r;
      if (py >= H) continue;
      const float wx = wxs[r], wy = wys[r], ex = 1.0f - wx, ey = 1.0f - wy;
      float gx = 0.f, gy = 0.f;
#pragma unroll
      for (int ch = 0; ch < CT; ++ch) {
        // a tap outside the image counts as value 0 (ATen within_bounds); offsets are always safe
        const float v00 = (tap[r].valid & 1) ? v[r][ch][0] : 0.f, v01 = (tap[r].valid & 2) ? v[r][ch][1] : 0.f;
        const float v10 = (tap[r].valid & 4) ? v[r][ch][2] : 0.f, v11 = (tap[r].valid & 8) ? v[r][ch][3] : 0.f;
        gx += g[r][ch] * ((v01 - v00) * ey + (v11 - v10) * wy);
        gy += g[r][ch] * ((v10 - v00) * ex + (v11 - v01) * wx);
      }
      // ATen: grad_grid = gix * (W-1)/2 * clipmask ; norm_grid backward: / (W-1) * 2
      float* dp = dflow + (long long)b * 2 * plane + (long long)py * W + px;
      __stcs(dp, __fdiv_rn(gx * (wm1 * 0.5f) * co[r].gmx, wm1) * 2.0f);
      __stcs(dp + plane, __fdiv_rn(gy * (hm1 * 0.5f) * co[r].gmy, hm1) * 2.0f);
    }
  } else {
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const int py = py0 + 8 * r;
    if (py >= H) continue;
    const long long pix = (long long)py * W;
    const Coord co = sample_point<BORDER>((float)px + fx[r], (float)py + fy[r], H, W);
    float wx, wy;
    const Tap tap = make_tap(co, H, W, wx, wy);
    const float ex = 1.0f - wx, ey = 1.0f - wy;
    float gx = 0.f, gy = 0.f;
    auto one = [&](int ch) {
      const float* xp = xb + ch * plane;
      const float g = __ldcs(gb + ch * plane + pix);
      // a tap outside the image counts as value 0 (ATen within_bounds); offsets are always safe
      float v00 = __ldg(xp + tap.o00), v01 = __ldg(xp + tap.o01), v10 = __ldg(xp + tap.o10), v11 = __ldg(xp + tap.o11);
      v00 = (tap.valid & 1) ? v00 : 0.f;
      v01 = (tap.valid & 2) ? v01 : 0.f;
      v10 = (tap.valid & 4) ? v10 : 0.f;
      v11 = (tap.valid & 8) ? v11 : 0.f;
      gx += g * ((v01 - v00) * ey + (v11 - v10) * wy);
      gy += g * ((v10 - v00) * ex + (v11 - v01) * wx);
      if (WITH_DX) {
        float* dp = dx + ((long long)b * C + ch) * plane;
        if (tap.w00 != 0.0f) atomicAdd(dp + tap.o00, g * tap.w00);
        if (tap.w01 != 0.0f) atomicAdd(dp + tap.o01, g * tap.w01);
        if (tap.w10 != 0.0f) atomicAdd(dp + tap.o10, g * tap.w10);
        if (tap.w11 != 0.0f) atomicAdd(dp + tap.o11, g * tap.w11);
      }
    };
    if constexpr (CT > 0) {
#pragma unroll
      for (int ch = 0; ch < CT; ++ch) one(ch);
    } else {
      for (int ch = 0; ch < C; ++ch) one(ch);
    }
    // ATen: grad_grid = gix * (W-1)/2 * clipmask ; norm_grid backward: / (W-1) * 2
    float* dp = dflow + (long long)b * 2 * plane + pix + px;
    __stcs(dp, __fdiv_rn(gx * (wm1 * 0.5f) * co.gmx, wm1) * 2.0f);
    __stcs(dp + plane, __fdiv_rn(gy * (hm1 * 0.5f) * co.gmy, hm1) * 2.0f);
  }
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Fast path for what the photometric loss calls (loss_flow.py:90-91): pad='border', C = 3, no image gradient.
// The generic kernels above cost ~225 SASS instructions per pixel (two IEEE divisions with slow-path branches,
// tap-validity masks, 64-bit address arithmetic) and were issue-bound at 40 % of HBM peak; this path needs ~90:
//  * a / (W-1) is computed as q = a*rc; q += fma(-q, W-1, a) * rc with rc = RN(1/(W-1)) -- the correctly rounded
//    quotient (checked against IEEE division on 4M samples per size, incl. W-1 = 255/511/1023), so the reference's
//    normalise -> un-normalise round trip is still reproduced;
//  * after the border clamp every tap is inside the image except x1 = W (y1 = H), which only occurs with weight
//    exactly 0: the index is clamped instead of masked;
//  * in-plane offsets are 32-bit.
// Persistent CTAs loop over 32 x 32 pixel tiles (rows of a thread 8 apart, lanes on consecutive x).  The flow of the
// NEXT tile is requested right after the gathers of the current tile have been issued, so the dependent chain
// flow -> address -> gather of one tile overlaps the gather latency of the previous one (ncu r1h: the single-shot
// version was latency-bound: long-scoreboard stalls 6.0 per issue, 21 resident warps per SM, DRAM 40 %).
template <bool BWD>
__global__ void __launch_bounds__(256)
flow_warp_border3_kernel(const float* __restrict__ x, const float* __restrict__ flow, const float* __restrict__ dout,
                         float* __restrict__ out, int H, int W, long long fsb, long long fsc, int tiles_x, int tiles_y,
                         int n_tiles, float rcw, float rch) {
  const unsigned plane = (unsigned)(H * W);
  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  int tile = blockIdx.x;
  if (tile >= n_tiles) return;

  int b, px, py0;
  unsigned pix[ROWS];
  float fx[ROWS], fy[ROWS];
  auto load_flow = [&](int t) {
    const int tx = t % tiles_x; t /= tiles_x;
    const int ty = t % tiles_y;
    b = t / tiles_y;
    px = min(tx * 32 + lx, W - 1);                   // columns past the end re-read the last column (stores masked)
    py0 = ty * (8 * ROWS) + ly;
    const float* fb = opaque(flow + (long long)b * fsb);
    const float* fb2 = opaque(fb + fsc);
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      pix[r] = (unsigned)(min(py0 + 8 * r, H - 1) * W + px);   // rows past the end re-read the last row
      fx[r] = __ldcs(fb + pix[r]);
      fy[r] = __ldcs(fb2 + pix[r]);
    }
    return tx * 32 + lx < W;
  };
  bool col_ok = load_flow(tile);

  while (true) {
    const int cb = b, cpy0 = py0;
    const bool c_ok = col_ok;
    unsigned cpix[ROWS];
    FastCoord c[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      cpix[r] = pix[r];
      c[r] = fast_coord((float)px + fx[r], (float)min(py0 + 8 * r, H - 1) + fy[r], H, W, wm1, hm1, rcw, rch);
    }
    // one opaque base pointer per channel plane: every tap pair is then a single IMAD.WIDE (offset * 4 + base)
    const float* xc[3];
    xc[0] = opaque(x + (long long)cb * 3 * plane);
    xc[1] = opaque(xc[0] + plane);
    xc[2] = opaque(xc[1] + plane);
    float v[ROWS][3][4];
    float g[ROWS][3];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const unsigned o0 = c[r].off, o1 = c[r].off + (unsigned)W;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {                // all 48 gathers in flight before any use
        const float* p0 = xc[ch] + o0;
        const float* p1 = xc[ch] + o1;
        v[r][ch][0] = __ldg(p0); v[r][ch][1] = __ldg(p0 + 1);
        v[r][ch][2] = __ldg(p1); v[r][ch][3] = __ldg(p1 + 1);
      }
    }
    if (BWD) {
      const float* gc[3];
      gc[0] = opaque(dout + (long long)cb * 3 * plane);
      gc[1] = opaque(gc[0] + plane);
      gc[2] = opaque(gc[1] + plane);
#pragma unroll
      for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) g[r][ch] = __ldcs(gc[ch] + cpix[r]);
    }
    const int next = tile + gridDim.x;
    const bool has_next = next < n_tiles;
    if (has_next) col_ok = load_flow(next);            // in flight while this tile is finished below

    if (!BWD) {
      float* oc[3];
      oc[0] = opaque(out + (long long)cb * 3 * plane);
      oc[1] = opaque(oc[0] + plane);
      oc[2] = opaque(oc[1] + plane);
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        if (c_ok && cpy0 + 8 * r < H) {
          const float wx = c[r].wx, wy = c[r].wy;
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            const float top = fmaf(wx, v[r][ch][1] - v[r][ch][0], v[r][ch][0]);
            const float bot = fmaf(wx, v[r][ch][3] - v[r][ch][2], v[r][ch][2]);
            __stcs(oc[ch] + cpix[r], fmaf(wy, bot - top, top));
          }
        }
      }
    } else {
      float* db = opaque(out + (long long)cb * 2 * plane);   // out = dflow [B,2,H,W]
      float* db2 = opaque(db + plane);
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        if (c_ok && cpy0 + 8 * r < H) {
          const float wx = c[r].wx, wy = c[r].wy, ex = 1.0f - wx, ey = 1.0f - wy;
          float gx = 0.f, gy = 0.f;
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            gx = fmaf(g[r][ch], (v[r][ch][1] - v[r][ch][0]) * ey + (v[r][ch][3] - v[r][ch][2]) * wy, gx);
            gy = fmaf(g[r][ch], (v[r][ch][2] - v[r][ch][0]) * ex + (v[r][ch][3] - v[r][ch][1]) * wx, gy);
          }
          // ATen: grad_grid = gix * (W-1)/2 * clipmask ; norm_grid backward: / (W-1) * 2
          __stcs(db + cpix[r], div_rn(gx * (wm1 * 0.5f) * c[r].gmx, wm1, rcw) * 2.0f);
          __stcs(db2 + cpix[r], div_rn(gy * (hm1 * 0.5f) * c[r].gmy, hm1, rch) * 2.0f);
        }
      }
    }
    if (!has_next) break;
    tile = next;
  }
}

}  // namespace

extern "C" int emip_flow_warp_fwd(const float* x, const float* flow, float* out, int B, int C, int H, int W,
                                  long long flow_stride_b, long long flow_stride_c, int pad_mode, void* stream) {
  return emip_flow_warp_fwd_ex(x, flow, out, B, C, H, W, flow_stride_b, flow_stride_c, pad_mode, EMIP_WARP_KERNEL_AUTO, nullptr,
                               nullptr, 0u, stream);
}

extern "C" int emip_flow_warp_fwd_ex(const float* x, const float* flow, float* out, int B, int C, int H, int W,
                                     long long flow_stride_b, long long flow_stride_c, int pad_mode, int kernel,
                                     unsigned* dev_stats, unsigned* host_stats, unsigned seq, void* stream) {
  if (B == 0) return EMIP_OK;
  EMIP_CHECK_ARG(kernel >= 0 && (kernel & 0xff) <= EMIP_WARP_KERNEL_STAGED, "flow_warp_fwd: bad kernel selector %d", kernel);
  const int ksel = kernel & 0xff, ctas = (kernel >> 8) == 1 ? 1 : 2;     // bits 8..: persistent CTAs per SM (tools; default 2)
  EMIP_CHECK_ARG(x && flow && out, "flow_warp_fwd: null pointer");
  EMIP_CHECK_ARG(B >= 0 && C > 0 && H > 1 && W > 1, "flow_warp_fwd: bad shape B=%d C=%d H=%d W=%d", B, C, H, W);
  EMIP_CHECK_ARG(pad_mode == EMIP_PAD_BORDER || pad_mode == EMIP_PAD_ZEROS, "flow_warp_fwd: bad pad_mode %d", pad_mode);
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles_x = (W + 31) / 32, tiles_y = (H + 8 * ROWS - 1) / (8 * ROWS);
  const long long nblk = (long long)B * tiles_x * tiles_y;
  EMIP_CHECK_ARG(nblk < 0x7fffffffLL, "flow_warp_fwd: problem too large");
#define LAUNCH(BD, CT) \
  flow_warp_fwd_kernel<BD, CT><<<(unsigned)nblk, 256, 0, st>>>(x, flow, out, C, H, W, flow_stride_b, flow_stride_c, tiles_x, tiles_y)
  const bool border = pad_mode == EMIP_PAD_BORDER;
  if (C == 3 && border && (long long)H * W * 4 < 0x7fffffffLL) {
    const float rcw = 1.0f / (float)(W - 1), rch = 1.0f / (float)(H - 1);
    const int grid = (int)(nblk < (long long)emip_num_sms() * ctas ? nblk : (long long)emip_num_sms() * ctas);
    int rc = EMIP_ENOSYS;
    if (ksel != EMIP_WARP_KERNEL_DIRECT)
      rc = flow_warp_staged_launch(false, x, flow, nullptr, out, B, H, W, flow_stride_b, flow_stride_c, ctas, dev_stats, host_stats, seq, st);
    if (rc != EMIP_OK && rc != EMIP_ENOSYS) return rc;
    if (rc == EMIP_ENOSYS && ksel == EMIP_WARP_KERNEL_STAGED) { emip_set_error("flow_warp_fwd: the staged kernel does not cover this shape / alignment"); return rc; }
    if (rc == EMIP_ENOSYS)
    flow_warp_border3_kernel<false><<<grid, 256, 0, st>>>(x, flow, nullptr, out, H, W, flow_stride_b, flow_stride_c,
                                                         tiles_x, tiles_y, (int)nblk, rcw, rch);
  } else if (C == 3) { if (border) LAUNCH(true, 3); else LAUNCH(false, 3); }
  else if (C == 2) { if (border) LAUNCH(true, 2); else LAUNCH(false, 2); }
  else { if (border) LAUNCH(true, 0); else LAUNCH(false, 0); }
#undef LAUNCH
  EMIP_CHECK_LAUNCH("flow_warp_fwd");
  return EMIP_OK;
}

extern "C" int emip_flow_warp_bwd(const float* x, const float* flow, const float* dout, float* dflow, float* dx,
                                  int B, int C, int H, int W, long long flow_stride_b, long long flow_stride_c,
                                  int pad_mode, void* stream) {
  return emip_flow_warp_bwd_ex(x, flow, dout, dflow, dx, B, C, H, W, flow_stride_b, flow_stride_c, pad_mode, EMIP_WARP_KERNEL_AUTO,
                               nullptr, nullptr, 0u, stream);
}

extern "C" int emip_flow_warp_bwd_ex(const float* x, const float* flow, const float* dout, float* dflow, float* dx,
                                     int B, int C, int H, int W, long long flow_stride_b, long long flow_stride_c,
                                     int pad_mode, int kernel, unsigned* dev_stats, unsigned* host_stats, unsigned seq,
                                     void* stream) {
  if (B == 0) return EMIP_OK;
  EMIP_CHECK_ARG(kernel >= 0 && (kernel & 0xff) <= EMIP_WARP_KERNEL_STAGED, "flow_warp_bwd: bad kernel selector %d", kernel);
  const int ksel = kernel & 0xff, ctas = (kernel >> 8) == 1 ? 1 : 2;
  EMIP_CHECK_ARG(x && flow && dout && dflow, "flow_warp_bwd: null pointer");
  EMIP_CHECK_ARG(B >= 0 && C > 0 && H > 1 && W > 1, "flow_warp_bwd: bad shape B=%d C=%d H=%d W=%d", B, C, H, W);
  EMIP_CHECK_ARG(pad_mode == EMIP_PAD_BORDER || pad_mode == EMIP_PAD_ZEROS, "flow_warp_bwd: bad pad_mode %d", pad_mode);
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles_x = (W + 31) / 32, tiles_y = (H + 8 * ROWS - 1) / (8 * ROWS);
  const long long nblk = (long long)B * tiles_x * tiles_y;
  EMIP_CHECK_ARG(nblk < 0x7fffffffLL, "flow_warp_bwd: problem too large");
#define LAUNCH(BD, DX, CT)                                                                                      \
  flow_warp_bwd_kernel<BD, DX, CT><<<(unsigned)nblk, 256, 0, st>>>(x, flow, dout, dflow, dx, C, H, W, flow_stride_b, \
                                                                   flow_stride_c, tiles_x, tiles_y)
  const bool border = pad_mode == EMIP_PAD_BORDER;
  if (C == 3 && border && dx == nullptr && (long long)H * W * 4 < 0x7fffffffLL) {
    const float rcw = 1.0f / (float)(W - 1), rch = 1.0f / (float)(H - 1);
    const int grid = (int)(nblk < (long long)emip_num_sms() * ctas ? nblk : (long long)emip_num_sms() * ctas);
    int rc = EMIP_ENOSYS;
    if (ksel != EMIP_WARP_KERNEL_DIRECT)
      rc = flow_warp_staged_launch(true, x, flow, dout, dflow, B, H, W, flow_stride_b, flow_stride_c, ctas, dev_stats, host_stats, seq, st);
    if (rc != EMIP_OK && rc != EMIP_ENOSYS) return rc;
    if (rc == EMIP_ENOSYS && ksel == EMIP_WARP_KERNEL_STAGED) { emip_set_error("flow_warp_bwd: the staged kernel does not cover this shape / alignment"); return rc; }
    if (rc == EMIP_ENOSYS)
    flow_warp_border3_kernel<true><<<grid, 256, 0, st>>>(x, flow, dout, dflow, H, W, flow_stride_b, flow_stride_c, tiles_x,
                                                        tiles_y, (int)nblk, rcw, rch);
  } else if (C == 3) {
    if (border) { if (dx) LAUNCH(true, true, 3); else LAUNCH(true, false, 3); }
    else { if (dx) LAUNCH(false, true, 3); else LAUNCH(false, false, 3); }
  } else {
    if (border) { if (dx) LAUNCH(true, true, 0); else LAUNCH(true, false, 0); }
    else { if (dx) LAUNCH(false, true, 0); else LAUNCH(false, false, 0); }
  }
#undef LAUNCH
  EMIP_CHECK_LAUNCH("flow_warp_bwd");
  return EMIP_OK;
}
