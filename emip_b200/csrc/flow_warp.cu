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

template <int VEC>
struct VecT;
template <>
struct VecT<1> { using type = float; };
template <>
struct VecT<2> { using type = float2; };
template <>
struct VecT<4> { using type = float4; };

template <int VEC>
__device__ __forceinline__ void load_vec(const float* p, float (&v)[VEC]) {
  using T = typename VecT<VEC>::type;
  T t = __ldg(reinterpret_cast<const T*>(p));
  const float* f = reinterpret_cast<const float*>(&t);
#pragma unroll
  for (int i = 0; i < VEC; ++i) v[i] = f[i];
}
template <int VEC>
__device__ __forceinline__ void store_vec_stream(float* p, const float (&v)[VEC]) {
  using T = typename VecT<VEC>::type;
  T t;
  float* f = reinterpret_cast<float*>(&t);
#pragma unroll
  for (int i = 0; i < VEC; ++i) f[i] = v[i];
  __stcs(reinterpret_cast<T*>(p), t);   // streaming store: output is not re-read by this kernel
}

// One thread = VEC horizontally adjacent output pixels, all C channels.
template <int VEC, bool BORDER>
__global__ void __launch_bounds__(256)
flow_warp_fwd_kernel(const float* __restrict__ x, const float* __restrict__ flow, float* __restrict__ out,
                     int B, int C, int H, int W, long long fsb, long long fsc) {
  const int wv = W / VEC;
  const long long total = (long long)B * H * wv;
  const long long plane = (long long)H * W;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int xv = (int)(idx % wv);
    long long r = idx / wv;
    int y = (int)(r % H);
    int b = (int)(r / H);
    int x0 = xv * VEC;
    const float* fp = flow + (long long)b * fsb + (long long)y * W + x0;
    float fx[VEC], fy[VEC];
    load_vec<VEC>(fp, fx);
    load_vec<VEC>(fp + fsc, fy);
    Tap tap[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      Coord c = sample_point<BORDER>((float)(x0 + i) + fx[i], (float)y + fy[i], H, W);
      float wx, wy;
      tap[i] = make_tap(c, H, W, wx, wy);
    }
    const float* xb = x + (long long)b * C * plane;
    float* ob = out + (long long)b * C * plane + (long long)y * W + x0;
    for (int ch = 0; ch < C; ++ch) {
      const float* xp = xb + ch * plane;
      float o[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        float v00 = __ldg(xp + tap[i].o00), v01 = __ldg(xp + tap[i].o01);
        float v10 = __ldg(xp + tap[i].o10), v11 = __ldg(xp + tap[i].o11);
        o[i] = v00 * tap[i].w00 + v01 * tap[i].w01 + v10 * tap[i].w10 + v11 * tap[i].w11;
      }
      store_vec_stream<VEC>(ob + ch * plane, o);
    }
  }
}

// Backward: dflow per pixel (no atomics); optional dx by atomic scatter.
template <int VEC, bool BORDER, bool WITH_DX>
__global__ void __launch_bounds__(256)
flow_warp_bwd_kernel(const float* __restrict__ x, const float* __restrict__ flow, const float* __restrict__ dout,
                     float* __restrict__ dflow, float* __restrict__ dx,
                     int B, int C, int H, int W, long long fsb, long long fsc) {
  const int wv = W / VEC;
  const long long total = (long long)B * H * wv;
  const long long plane = (long long)H * W;
  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int xv = (int)(idx % wv);
    long long r = idx / wv;
    int y = (int)(r % H);
    int b = (int)(r / H);
    int x0 = xv * VEC;
    const float* fp = flow + (long long)b * fsb + (long long)y * W + x0;
    float fx[VEC], fy[VEC];
    load_vec<VEC>(fp, fx);
    load_vec<VEC>(fp + fsc, fy);
    Tap tap[VEC];
    Coord co[VEC];
    float wx[VEC], wy[VEC], gx[VEC], gy[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      co[i] = sample_point<BORDER>((float)(x0 + i) + fx[i], (float)y + fy[i], H, W);
      tap[i] = make_tap(co[i], H, W, wx[i], wy[i]);
      gx[i] = 0.0f;
      gy[i] = 0.0f;
    }
    const float* xb = x + (long long)b * C * plane;
    const float* gb = dout + (long long)b * C * plane + (long long)y * W + x0;
    for (int ch = 0; ch < C; ++ch) {
      const float* xp = xb + ch * plane;
      float g[VEC];
      load_vec<VEC>(gb + ch * plane, g);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        // a tap outside the image counts as value 0 (ATen within_bounds); offsets are always safe
        float v00 = (tap[i].valid & 1) ? __ldg(xp + tap[i].o00) : 0.0f;
        float v01 = (tap[i].valid & 2) ? __ldg(xp + tap[i].o01) : 0.0f;
        float v10 = (tap[i].valid & 4) ? __ldg(xp + tap[i].o10) : 0.0f;
        float v11 = (tap[i].valid & 8) ? __ldg(xp + tap[i].o11) : 0.0f;
        float ex = 1.0f - wx[i], ey = 1.0f - wy[i];
        gx[i] += g[i] * ((v01 - v00) * ey + (v11 - v10) * wy[i]);
        gy[i] += g[i] * ((v10 - v00) * ex + (v11 - v01) * wx[i]);
        if (WITH_DX) {
          float* dp = dx + ((long long)b * C + ch) * plane;
          if (tap[i].w00 != 0.0f) atomicAdd(dp + tap[i].o00, g[i] * tap[i].w00);
          if (tap[i].w01 != 0.0f) atomicAdd(dp + tap[i].o01, g[i] * tap[i].w01);
          if (tap[i].w10 != 0.0f) atomicAdd(dp + tap[i].o10, g[i] * tap[i].w10);
          if (tap[i].w11 != 0.0f) atomicAdd(dp + tap[i].o11, g[i] * tap[i].w11);
        }
      }
    }
    float ox[VEC], oy[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      // ATen: grad_grid = gix * (W-1)/2 * clipmask ; norm_grid backward: / (W-1) * 2
      ox[i] = __fdiv_rn(gx[i] * (wm1 * 0.5f) * co[i].gmx, wm1) * 2.0f;
      oy[i] = __fdiv_rn(gy[i] * (hm1 * 0.5f) * co[i].gmy, hm1) * 2.0f;
    }
    float* dp = dflow + (long long)b * 2 * plane + (long long)y * W + x0;
    store_vec_stream<VEC>(dp, ox);
    store_vec_stream<VEC>(dp + plane, oy);
  }
}

int pick_vec(const void* a, const void* b, const void* c, int W, long long fsb, long long fsc) {
  auto al = [](const void* p, int bytes) { return (reinterpret_cast<uintptr_t>(p) % bytes) == 0; };
  if (W % 4 == 0 && fsb % 4 == 0 && fsc % 4 == 0 && al(a, 16) && al(b, 16) && al(c, 16)) return 4;
  if (W % 2 == 0 && fsb % 2 == 0 && fsc % 2 == 0 && al(a, 8) && al(b, 8) && al(c, 8)) return 2;
  return 1;
}

int grid_for(long long total_threads) {
  long long blocks = (total_threads + 255) / 256;
  long long cap = (long long)emip_num_sms() * 32;   // grid-stride above 32 CTAs/SM
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

extern "C" int emip_flow_warp_fwd(const float* x, const float* flow, float* out, int B, int C, int H, int W,
                                  long long flow_stride_b, long long flow_stride_c, int pad_mode, void* stream) {
  if (B == 0) return EMIP_OK;
  EMIP_CHECK_ARG(x && flow && out, "flow_warp_fwd: null pointer");
  EMIP_CHECK_ARG(B >= 0 && C > 0 && H > 1 && W > 1, "flow_warp_fwd: bad shape B=%d C=%d H=%d W=%d", B, C, H, W);
  EMIP_CHECK_ARG(pad_mode == EMIP_PAD_BORDER || pad_mode == EMIP_PAD_ZEROS, "flow_warp_fwd: bad pad_mode %d", pad_mode);
  if (B == 0) return EMIP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int vec = pick_vec(flow, out, flow + flow_stride_c, W, flow_stride_b, flow_stride_c);
  long long total = (long long)B * H * (W / vec);
  int grid = grid_for(total);
#define LAUNCH(V, BD) flow_warp_fwd_kernel<V, BD><<<grid, 256, 0, st>>>(x, flow, out, B, C, H, W, flow_stride_b, flow_stride_c)
  bool border = pad_mode == EMIP_PAD_BORDER;
  if (vec == 4) { if (border) LAUNCH(4, true); else LAUNCH(4, false); }
  else if (vec == 2) { if (border) LAUNCH(2, true); else LAUNCH(2, false); }
  else { if (border) LAUNCH(1, true); else LAUNCH(1, false); }
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
  if (B == 0) return EMIP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int vec = pick_vec(flow, dout, flow + flow_stride_c, W, flow_stride_b, flow_stride_c);
  if (vec > 1 && (reinterpret_cast<uintptr_t>(dflow) % (4 * vec)) != 0) vec = 1;
  long long total = (long long)B * H * (W / vec);
  int grid = grid_for(total);
#define LAUNCH(V, BD, DX) \
  flow_warp_bwd_kernel<V, BD, DX><<<grid, 256, 0, st>>>(x, flow, dout, dflow, dx, B, C, H, W, flow_stride_b, flow_stride_c)
#define LAUNCH_V(V)                                                  \
  do {                                                               \
    if (border) { if (dx) LAUNCH(V, true, true); else LAUNCH(V, true, false); }   \
    else { if (dx) LAUNCH(V, false, true); else LAUNCH(V, false, false); }        \
  } while (0)
  bool border = pad_mode == EMIP_PAD_BORDER;
  if (vec == 4) LAUNCH_V(4); else if (vec == 2) LAUNCH_V(2); else LAUNCH_V(1);
#undef LAUNCH_V
#undef LAUNCH
  EMIP_CHECK_LAUNCH("flow_warp_bwd");
  return EMIP_OK;
}
