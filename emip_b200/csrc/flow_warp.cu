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

// Forward.  CTA = 32 x 8 output pixels of one sample; one thread per pixel, all C channels.
// Consecutive lanes = consecutive pixels, so for the piecewise-smooth flows the model produces the four
// tap requests of a warp touch 1-2 cache lines each (ncu r1a: the earlier 4-pixels-per-thread mapping made
// every tap request span ~30 sectors and held the kernel to 19 % of HBM peak).  The 8-row tile lets the
// bottom taps of row y and the top taps of row y+1 hit the same L1 lines.
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
  const int py = ty * 8 + (threadIdx.x >> 5);
  if (px >= W || py >= H) return;
  const long long pix = (long long)py * W + px;
  const float* fp = flow + (long long)b * fsb + pix;
  const float fx = __ldcs(fp), fy = __ldcs(fp + fsc);          // streamed: read exactly once
  Coord c = sample_point<BORDER>((float)px + fx, (float)py + fy, H, W);
  float wx, wy;
  const Tap tap = make_tap(c, H, W, wx, wy);
  const float* xb = x + (long long)b * C * plane;
  float* ob = out + (long long)b * C * plane + pix;
  if constexpr (CT > 0) {
    float v[CT > 0 ? CT : 1][4];
#pragma unroll
    for (int ch = 0; ch < CT; ++ch) {                           // all 4*C gathers in flight before any use
      const float* xp = xb + ch * plane;
      v[ch][0] = __ldg(xp + tap.o00); v[ch][1] = __ldg(xp + tap.o01);
      v[ch][2] = __ldg(xp + tap.o10); v[ch][3] = __ldg(xp + tap.o11);
    }
#pragma unroll
    for (int ch = 0; ch < CT; ++ch)
      __stcs(ob + ch * plane, v[ch][0] * tap.w00 + v[ch][1] * tap.w01 + v[ch][2] * tap.w10 + v[ch][3] * tap.w11);
  } else {
    for (int ch = 0; ch < C; ++ch) {
      const float* xp = xb + ch * plane;
      float v00 = __ldg(xp + tap.o00), v01 = __ldg(xp + tap.o01), v10 = __ldg(xp + tap.o10), v11 = __ldg(xp + tap.o11);
      __stcs(ob + ch * plane, v00 * tap.w00 + v01 * tap.w01 + v10 * tap.w10 + v11 * tap.w11);
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
  const int py = ty * 8 + (threadIdx.x >> 5);
  if (px >= W || py >= H) return;
  const long long pix = (long long)py * W + px;
  const float* fp = flow + (long long)b * fsb + pix;
  const float fx = __ldcs(fp), fy = __ldcs(fp + fsc);
  const Coord co = sample_point<BORDER>((float)px + fx, (float)py + fy, H, W);
  float wx, wy;
  const Tap tap = make_tap(co, H, W, wx, wy);
  const float ex = 1.0f - wx, ey = 1.0f - wy;
  const float* xb = x + (long long)b * C * plane;
  const float* gb = dout + (long long)b * C * plane + pix;
  float gx = 0.f, gy = 0.f;
  auto one = [&](int ch) {
    const float* xp = xb + ch * plane;
    const float g = __ldcs(gb + ch * plane);
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
  float* dp = dflow + (long long)b * 2 * plane + pix;
  __stcs(dp, __fdiv_rn(gx * (wm1 * 0.5f) * co.gmx, wm1) * 2.0f);
  __stcs(dp + plane, __fdiv_rn(gy * (hm1 * 0.5f) * co.gmy, hm1) * 2.0f);
}

}  // namespace

extern "C" int emip_flow_warp_fwd(const float* x, const float* flow, float* out, int B, int C, int H, int W,
                                  long long flow_stride_b, long long flow_stride_c, int pad_mode, void* stream) {
  if (B == 0) return EMIP_OK;
  EMIP_CHECK_ARG(x && flow && out, "flow_warp_fwd: null pointer");
  EMIP_CHECK_ARG(B >= 0 && C > 0 && H > 1 && W > 1, "flow_warp_fwd: bad shape B=%d C=%d H=%d W=%d", B, C, H, W);
  EMIP_CHECK_ARG(pad_mode == EMIP_PAD_BORDER || pad_mode == EMIP_PAD_ZEROS, "flow_warp_fwd: bad pad_mode %d", pad_mode);
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles_x = (W + 31) / 32, tiles_y = (H + 7) / 8;
  const long long nblk = (long long)B * tiles_x * tiles_y;
  EMIP_CHECK_ARG(nblk < 0x7fffffffLL, "flow_warp_fwd: problem too large");
#define LAUNCH(BD, CT) \
  flow_warp_fwd_kernel<BD, CT><<<(unsigned)nblk, 256, 0, st>>>(x, flow, out, C, H, W, flow_stride_b, flow_stride_c, tiles_x, tiles_y)
  const bool border = pad_mode == EMIP_PAD_BORDER;
  if (C == 3) { if (border) LAUNCH(true, 3); else LAUNCH(false, 3); }
  else if (C == 2) { if (border) LAUNCH(true, 2); else LAUNCH(false, 2); }
  else { if (border) LAUNCH(true, 0); else LAUNCH(false, 0); }
#undef LAUNCH
  EMIP_CHECK_LAUNCH("flow_warp_fwd");
  return EMIP_OK;
}

extern "C" int emip_flow_warp_bwd(const float* x, const float* flow, const float* dout, float* dflow, float* dx,
                                  int B, int C, int H, int W, long long flow_stride_b, long long flow_stride_c,
                                  int pad_mode, void* stream) {
  if (B == 0) return EMIP_OK;
  EMIP_CHECK_ARG(x && flow && dout && dflow, "flow_warp_bwd: null pointer");
  EMIP_CHECK_ARG(B >= 0 && C > 0 && H > 1 && W > 1, "flow_warp_bwd: bad shape B=%d C=%d H=%d W=%d", B, C, H, W);
  EMIP_CHECK_ARG(pad_mode == EMIP_PAD_BORDER || pad_mode == EMIP_PAD_ZEROS, "flow_warp_bwd: bad pad_mode %d", pad_mode);
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles_x = (W + 31) / 32, tiles_y = (H + 7) / 8;
  const long long nblk = (long long)B * tiles_x * tiles_y;
  EMIP_CHECK_ARG(nblk < 0x7fffffffLL, "flow_warp_bwd: problem too large");
#define LAUNCH(BD, DX, CT)                                                                                      \
  flow_warp_bwd_kernel<BD, DX, CT><<<(unsigned)nblk, 256, 0, st>>>(x, flow, dout, dflow, dx, C, H, W, flow_stride_b, \
                                                                   flow_stride_c, tiles_x, tiles_y)
  const bool border = pad_mode == EMIP_PAD_BORDER;
  if (C == 3) {
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
