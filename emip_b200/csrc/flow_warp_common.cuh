// Helpers shared by the pad='border', C = 3 fast paths of flow_warp (flow_warp.cu, flow_warp_staged.cu).
#pragma once
#include "common.cuh"

namespace k3 {

constexpr int ROWS = 4;   // pixels per thread (rows 8 apart) in the 32 x 32 pixel tile of a 256-thread CTA

struct FastCoord {
  unsigned off;          // in-plane offset of the top-left tap (y0 * W + x0); the 2x2 block is off, off+1, off+W, off+W+1
  float wx, wy, gmx, gmy;
};

__device__ __forceinline__ float div_rn(float a, float b, float rc) {
  float q = a * rc;
  return fmaf(fmaf(-q, b, a), rc, q);
}

struct FastCoordXY {
  int x0, y0;            // top-left tap, 0 <= x0 <= W-2, 0 <= y0 <= H-2
  float wx, wy, gmx, gmy;
};

// x0 is clamped to W-2 (y0 to H-2) so that the 2x2 tap block is always inside the image: at ix = W-1 this gives
// (x0, wx) = (W-2, 1) instead of the reference's (W-1, 0) -- the same interpolated value and the same gradients.
__device__ __forceinline__ FastCoordXY fast_coord_xy(float u, float v, int H, int W, float wm1, float hm1, float rcw,
                                                     float rch) {
  // warp_utils.py:21-22 then ATen grid_sampler_unnormalize (align_corners=True)
  float ix = (((div_rn(2.0f * u, wm1, rcw) - 1.0f) + 1.0f) * 0.5f) * wm1;
  float iy = (((div_rn(2.0f * v, hm1, rch) - 1.0f) + 1.0f) * 0.5f) * hm1;
  FastCoordXY c;
  c.gmx = (ix > 0.0f && ix < wm1) ? 1.0f : 0.0f;     // ATen clip_coordinates_set_grad
  c.gmy = (iy > 0.0f && iy < hm1) ? 1.0f : 0.0f;
  ix = fminf(fmaxf(ix, 0.0f), wm1);
  iy = fminf(fmaxf(iy, 0.0f), hm1);
  c.x0 = min(__float2int_rd(ix), W - 2);
  c.y0 = min(__float2int_rd(iy), H - 2);
  c.wx = ix - (float)c.x0;
  c.wy = iy - (float)c.y0;
  return c;
}

__device__ __forceinline__ FastCoord fast_coord(float u, float v, int H, int W, float wm1, float hm1, float rcw,
                                                float rch) {
  const FastCoordXY q = fast_coord_xy(u, v, H, W, wm1, hm1, rcw, rch);
  FastCoord c;
  c.off = (unsigned)(q.y0 * W + q.x0);
  c.wx = q.wx; c.wy = q.wy; c.gmx = q.gmx; c.gmy = q.gmy;
  return c;
}

// Makes a per-sample base pointer opaque to the optimiser: otherwise nvcc folds the 64-bit batch offset into every
// tap index and spends four instructions per address (IMAD.WIDE + IADD3 + LEA + LEA.HI.X) instead of one IMAD.WIDE.
template <typename T>
__device__ __forceinline__ T* opaque(T* p) {
  asm volatile("" : "+l"(p));
  return p;
}


}  // namespace k3
