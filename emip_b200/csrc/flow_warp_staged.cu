// K3, shared-memory-staged variant of the photometric-loss warp (pad='border', C = 3, no image gradient).
//
// Replaces reference loss/warp_utils.py:83-93 for the call pattern of loss/loss_flow.py:90-91, like the direct-gather
// fast path in flow_warp.cu, but every global read goes through the TMA unit and the bilinear taps are read from
// shared memory.  A persistent CTA walks 32 x 32 pixel tiles with 8 worker warps and one producer warp:
//
//   producer (warp 8)  keeps a ring of flow tiles in flight (cp.async.bulk.tensor.4d over the strided [B,2,H,W] flow
//                      view, so the channel slices of loss_flow.py:90-91 need no copy); for each tile it combines the
//                      flow ranges the worker warps report into a conservative window of the image that contains
//                      every tap of the tile, and requests it (box 44x40 or 52x48 floats x 3 planes; the backward
//                      pass adds the 32x32x3 tile of the incoming gradient on the same mbarrier);
//   workers (warps 0-7) report the min / max flow of the NEXT tile (one redux.sync per bound), then read the flow of
//                      their 4 pixels of the current tile from shared memory, reproduce the reference's normalise ->
//                      un-normalise coordinate round trip, read the 2x2x3 taps with 12 LDS that share one address
//                      register, blend and stream the result out.
//
// Why: ncu (profiles/r1h_k3_*.txt) shows the direct-gather kernel limited by the L1 data pipe -- a 32-lane gather of
// 4-byte taps straddles two 128-byte lines, 2.6 wavefronts per request, l1tex__data_pipe_lsu_wavefronts 61 % at 63 us
// and ~82 % at 59 us -- not by HBM.  Shared memory has no line granularity and the windows arrive without occupying
// LSU issue slots.  A tile whose window does not fit 52 x 48 (flow range > ~12 px inside the tile) gathers from global
// memory for that tile only.  TMA facts measured on the way (tools/probe/tma_window_probe.cu): the x coordinate of a
// tiled fp32 box must be 16-byte aligned (else the kernel faults); a window costs ~0.8 us latency and the unit
// sustains one box per ~0.25 us per SM whatever its size.
#include <type_traits>
#include "common.cuh"
#include "tc_common.cuh"
#include "flow_warp_common.cuh"
#include "flow_warp_staged.cuh"

namespace {

using k3::ROWS;
using k3::div_rn;
using k3::opaque;

constexpr int NBOX = 2;
constexpr int WORKERS = 8;                                     // worker warps
constexpr int THREADS = (WORKERS + 1) * 32;
constexpr int NFLOW = 4;                                       // flow tiles in flight
constexpr int NWIN = 2;                                        // image windows (+ gradient tiles) in flight
constexpr int FLOW_BYTES = 2 * 32 * 32 * 4;                    // 8 KB
constexpr int GRAD_BYTES = 3 * 32 * 32 * 4;                    // 12 KB (backward only)

template <bool BWD> struct Layout {
  // window variants (floats); width * 4 % 16 == 0.  Forward: pitch 64 = bank-conflict-free taps for x-monotone rows.
  static constexpr int BOX_AW = BWD ? 44 : 64, BOX_AH = 40, BOX_BW = BWD ? 52 : 64, BOX_BH = 48;
  static constexpr int WIN_BYTES = BOX_BW * BOX_BH * 3 * 4;
  static_assert(WIN_BYTES % 128 == 0 && BOX_AW * BOX_AH * 12 <= WIN_BYTES, "window buffer");
  static constexpr int NFL = BWD ? 3 : NFLOW;
  static constexpr int OFF_WIN = 0;
  static constexpr int OFF_GRAD = OFF_WIN + NWIN * WIN_BYTES;
  static constexpr int OFF_FLOW = OFF_GRAD + (BWD ? NWIN * GRAD_BYTES : 0);
  static constexpr int OFF_BARS = OFF_FLOW + NFL * FLOW_BYTES;          // flow_full, flow_empty, win_full, win_empty, box_full
  static constexpr int OFF_HDR = (OFF_BARS + 8 * (2 * NFL + 3 * NWIN) + 15) / 16 * 16;
  static constexpr int OFF_BOX = OFF_HDR + NWIN * 16;                   // [NWIN][WORKERS] int4: flow min / max keys per warp
  static constexpr int BYTES = OFF_BOX + NWIN * WORKERS * 16 + 128 /*alignment slack*/;
  static_assert(OFF_HDR % 16 == 0, "header alignment");
  static_assert(2 * (BYTES + 1024) <= 228 * 1024, "two CTAs per SM");
};

struct StagedMaps { CUtensorMap win[NBOX]; CUtensorMap flow; CUtensorMap grad; };

struct StagedParams {
  const float* x;
  float* out;
  int H, W, tiles_x, tiles_y, n_tiles;
  int gx, gy, gb;          // gridDim.x decomposed into (samples, tile rows, tile columns): the persistent stride
  float rcw, rch;
  long long* prof;         // optional [gridDim.x][8] wait-cycle profile (tools/k3_roles.py), else null
  unsigned* dev_stats;     // [3]: tiles that gathered from global memory, CTAs finished, (unused); null = no feedback
  volatile unsigned* host_stats;   // mapped host memory [2]: {launch sequence number, per-mille of such tiles}
  unsigned seq;
};

// Bounded mbarrier wait: a protocol bug traps instead of hanging the GPU.
__device__ __forceinline__ long long wait_bar(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  unsigned spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"     // suspend-time hint (ns): fewer spin trips
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(2000u)
        : "memory");
    if (ok) break;
    if (++spins > (1u << 22)) {
      printf("emip flow_warp_staged: mbarrier timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar,
             parity);
      __trap();
    }
  }
  return clock64() - t0;
}

// 12 taps per pixel from a staged window of pitch PW, height PH.  rel = (y0 - ymin) * PW + (x0 - xmin) is one dp2a:
// lo16(xy) * 1 + hi16(xy) * PW - base.
template <int PW, int PH>
__device__ __forceinline__ void gather_staged(const float* __restrict__ sb, const unsigned (&xy)[ROWS], int xmin, int ymin,
                                              float (&v)[ROWS][3][4]) {
  const int nbase = -(ymin * PW + xmin);
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const float* p = sb + __dp2a_lo((int)xy[r], (PW << 8) | 1, nbase);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      v[r][ch][0] = p[ch * PW * PH];
      v[r][ch][1] = p[ch * PW * PH + 1];
      v[r][ch][2] = p[ch * PW * PH + PW];
      v[r][ch][3] = p[ch * PW * PH + PW + 1];
    }
  }
}

template <bool BWD>
__global__ void __launch_bounds__(THREADS, 2)
flow_warp_staged_kernel(const __grid_constant__ StagedMaps maps, const __grid_constant__ StagedParams P) {
  using L = Layout<BWD>;
  constexpr int NFL = L::NFL;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = (uint32_t)__cvta_generic_to_shared(smem_raw);
  unsigned char* sm = smem_raw + ((128u - (raw & 127u)) & 127u);
  const uint32_t sm0 = (uint32_t)__cvta_generic_to_shared(sm);
  const uint32_t flow_full = sm0 + L::OFF_BARS, flow_empty = flow_full + 8 * NFL, win_full = flow_empty + 8 * NFL,
                 win_empty = win_full + 8 * NWIN, box_full = win_empty + 8 * NWIN;
  int4* hdr = reinterpret_cast<int4*>(sm + L::OFF_HDR);                             // [NWIN] (xmin, ymin, variant, -)
  int4* boxes = reinterpret_cast<int4*>(sm + L::OFF_BOX);                           // [NWIN][WORKERS]

  const int H = P.H, W = P.W;
  const unsigned plane = (unsigned)(H * W);
  const int lx = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if ((int)blockIdx.x >= P.n_tiles) return;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NFL; ++i) { tc::mbar_init(flow_full + 8 * i, 1); tc::mbar_init(flow_empty + 8 * i, WORKERS); }
    for (int i = 0; i < NWIN; ++i) {
      tc::mbar_init(win_full + 8 * i, 1);
      tc::mbar_init(win_empty + 8 * i, WORKERS);
      tc::mbar_init(box_full + 8 * i, WORKERS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();                                       // the only CTA-wide barrier of the kernel

  // ---- tile walk shared by both roles: advanced by the grid stride without divisions
  const int my_tiles = (P.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  int tx, ty, tb;
  {
    int t = blockIdx.x;
    tx = t % P.tiles_x; t /= P.tiles_x;
    ty = t % P.tiles_y;
    tb = t / P.tiles_y;
  }
  auto advance = [&](int& ax, int& ay, int& ab) {
    ax += P.gx;
    if (ax >= P.tiles_x) { ax -= P.tiles_x; ++ay; }
    ay += P.gy;
    if (ay >= P.tiles_y) { ay -= P.tiles_y; ++ab; }
    ab += P.gb;
  };
  // order-preserving float <-> int keys: the warp-wide min / max then is one redux.sync
  auto key = [](float f) { const int i = __float_as_int(f); return i ^ ((i >> 31) & 0x7fffffff); };
  auto unkey = [](int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); };

  if (warp == WORKERS) {
    // =============================== producer ===============================
    // Keeps the flow ring full and turns the workers' per-warp flow ranges of tile m into the window request of tile m.
    int fx_t = tx, fy_t = ty, fb_t = tb;                 // tile whose flow is requested next
    int issued = 0;
    long long w_fe = 0, w_box = 0, w_we = 0;
    unsigned n_global = 0;                               // tiles whose window did not fit
    const long long t_start = clock64();
    auto request_flow = [&]() {                          // lane 0 only
      const int s = issued % NFL;
      w_fe += wait_bar(flow_empty + 8 * s, ((uint32_t)(issued / NFL) & 1u) ^ 1u);
      tc::mbar_expect_tx(flow_full + 8 * s, FLOW_BYTES);
      tc::tma_load_4d(sm0 + L::OFF_FLOW + s * FLOW_BYTES, &maps.flow, flow_full + 8 * s, fx_t * 32, fy_t * 32, 0, fb_t);
      advance(fx_t, fy_t, fb_t);
      ++issued;
    };
    if (lx == 0)
      for (int i = 0; i < NFL - 1 && i < my_tiles; ++i) request_flow();
    for (int m = 0; m < my_tiles; ++m) {
      const int ws = m % NWIN;
      w_box += wait_bar(box_full + 8 * ws, (uint32_t)(m / NWIN) & 1u);
      const int4 q = boxes[ws * WORKERS + (lx & (WORKERS - 1))];
      const float xlo = unkey(__reduce_min_sync(0xffffffffu, q.x)), ylo = unkey(__reduce_min_sync(0xffffffffu, q.y));
      const float xhi = unkey(__reduce_max_sync(0xffffffffu, q.z)), yhi = unkey(__reduce_max_sync(0xffffffffu, q.w));
      if (lx == 0) {
        // every tap of the tile lies in [xs, xe] x [ys, ye]: u = px + fx, the coordinate round trip moves it by far
        // less than one pixel, the border clamp pulls it into the image (x0 <= W-2, taps x0 and x0+1)
        const float x_lo = (float)(tx * 32) + xlo, x_hi = (float)(tx * 32 + 31) + xhi;
        const float y_lo = (float)(ty * 32) + ylo, y_hi = (float)(ty * 32 + 31) + yhi;
        const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
        int xs = (int)floorf(fminf(fmaxf(x_lo, 0.f), wm1)) - 1, xe = (int)floorf(fminf(fmaxf(x_hi, 0.f), wm1)) + 2;
        int ys = (int)floorf(fminf(fmaxf(y_lo, 0.f), hm1)) - 1, ye = (int)floorf(fminf(fmaxf(y_hi, 0.f), hm1)) + 2;
        xs = max(xs, 0) & ~3;                            // TMA: the window must start on a 16-byte boundary in x
        ys = max(ys, 0);
        xe = min(xe, W - 1);
        ye = min(ye, H - 1);
        const int nx = xe - xs + 1, ny = ye - ys + 1;
        const int variant = (nx <= L::BOX_AW && ny <= L::BOX_AH) ? 0 : (nx <= L::BOX_BW && ny <= L::BOX_BH) ? 1 : 2;
        n_global += variant == 2 ? 1u : 0u;
        w_we += wait_bar(win_empty + 8 * ws, ((uint32_t)(m / NWIN) & 1u) ^ 1u);
        hdr[ws] = make_int4(xs, ys, variant, 0);
        const uint32_t bar = win_full + 8 * ws, dst = sm0 + L::OFF_WIN + ws * L::WIN_BYTES;
        const uint32_t extra = BWD ? GRAD_BYTES : 0;
        if (variant == 0) {
          tc::mbar_expect_tx(bar, L::BOX_AW * L::BOX_AH * 12 + extra);
          tc::tma_load_3d(dst, &maps.win[0], bar, xs, ys, tb * 3);
        } else if (variant == 1) {
          tc::mbar_expect_tx(bar, L::BOX_BW * L::BOX_BH * 12 + extra);
          tc::tma_load_3d(dst, &maps.win[1], bar, xs, ys, tb * 3);
        } else if (BWD) {
          tc::mbar_expect_tx(bar, extra);                // window too large: this tile gathers from global memory
        } else {
          tc::mbar_arrive(bar);
        }
        if (BWD) tc::tma_load_3d(sm0 + L::OFF_GRAD + ws * GRAD_BYTES, &maps.grad, bar, tx * 32, ty * 32, tb * 3);
        if (issued < my_tiles) request_flow();           // tile m + NFL - 1 into the slot of tile m - 1
      }
      advance(tx, ty, tb);
    }
    if (P.prof != nullptr && lx == 0) {
      long long* q = P.prof + (size_t)blockIdx.x * 8;
      q[0] = clock64() - t_start; q[1] = w_fe; q[2] = w_box; q[3] = w_we;
    }
    // feedback for the launcher's kernel choice: the CTA that finishes last publishes the share of such tiles
    if (P.dev_stats != nullptr && lx == 0) {
      if (n_global) atomicAdd(P.dev_stats, n_global);
      __threadfence();
      if (atomicAdd(P.dev_stats + 1, 1u) == gridDim.x - 1) {
        __threadfence();
        const unsigned total = atomicExch(P.dev_stats, 0u);
        P.dev_stats[1] = 0u;
        P.host_stats[1] = (unsigned)(1000ull * total / (unsigned)P.n_tiles);
        __threadfence_system();
        P.host_stats[0] = P.seq;
      }
    }
    return;
  }

  // ================================= workers =================================
  const int ly = warp;
  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  const float hw = 0.5f * wm1, hh = 0.5f * hm1, rc2w = 2.0f * P.rcw, rc2h = 2.0f * P.rch;
  float* const o0 = opaque(P.out);
  float* const o1 = opaque(P.out + plane);
  float* const o2 = BWD ? nullptr : opaque(P.out + 2 * (size_t)plane);
  long long w_flow = 0, w_win = 0;
  const long long t_start = clock64();

  // flow range of my four pixels of tile m -> boxes[m % NWIN][warp]; the producer combines the eight ranges.
  // The slot is free: its previous content (tile m - 2) was consumed before the window of tile m - 2 was requested,
  // and this warp has gathered that window already.
  auto report = [&](int m) {
    const int fs = m % NFL;
    w_flow += wait_bar(flow_full + 8 * fs, (uint32_t)(m / NFL) & 1u);
    const float* fsm = reinterpret_cast<const float*>(sm + L::OFF_FLOW + fs * FLOW_BYTES) + ly * 32 + lx;
    float a[ROWS], b[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) { a[r] = fsm[r * 256]; b[r] = fsm[1024 + r * 256]; }
    const int kxlo = __reduce_min_sync(0xffffffffu, key(fminf(fminf(a[0], a[1]), fminf(a[2], a[3]))));
    const int kylo = __reduce_min_sync(0xffffffffu, key(fminf(fminf(b[0], b[1]), fminf(b[2], b[3]))));
    const int kxhi = __reduce_max_sync(0xffffffffu, key(fmaxf(fmaxf(a[0], a[1]), fmaxf(a[2], a[3]))));
    const int kyhi = __reduce_max_sync(0xffffffffu, key(fmaxf(fmaxf(b[0], b[1]), fmaxf(b[2], b[3]))));
    if (lx == 0) {
      boxes[(m % NWIN) * WORKERS + warp] = make_int4(kxlo, kylo, kxhi, kyhi);
      tc::mbar_arrive(box_full + 8 * (m % NWIN));
    }
  };
  report(0);

  for (int n = 0; n < my_tiles; ++n) {
    if (n + 1 < my_tiles) report(n + 1);                 // lets the producer request the next window while I gather

    // ---- flow of my four pixels (landed: seen by report(n)) -> tap coordinates
    const int fs = n % NFL;
    const int px = tx * 32 + lx, py0 = ty * (8 * ROWS) + ly;
    const float* fsm = reinterpret_cast<const float*>(sm + L::OFF_FLOW + fs * FLOW_BYTES) + ly * 32 + lx;
    float fx[ROWS], fy[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) { fx[r] = fsm[r * 256]; fy[r] = fsm[1024 + r * 256]; }
    unsigned xy[ROWS];       // x0 | y0 << 16
    unsigned gm = 0;         // BWD: bit 2r = ix strictly inside (0, W-1), bit 2r+1 = iy strictly inside (0, H-1)
    float wxs[ROWS], wys[ROWS];
    const float fpx = (float)px, fpy = (float)py0;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      // pixels past the image edge read flow 0 (TMA zero fill) and are masked below
      const float u = fpx + fx[r], v = (fpy + (float)(8 * r)) + fy[r];
      // warp_utils.py:21-22 then ATen grid_sampler_unnormalize (align_corners=True), as k3::fast_coord_xy
      // 2u / (W-1) is u / ((W-1)/2) and (t * 0.5) * (W-1) is t * ((W-1)/2): the same real numbers, rounded once, so the
      // two multiplications by a power of two are folded into the constants (hw = (W-1)/2 and RN(1/hw) = 2 RN(1/(W-1)))
      float ix = ((div_rn(u, hw, rc2w) - 1.0f) + 1.0f) * hw;
      float iy = ((div_rn(v, hh, rc2h) - 1.0f) + 1.0f) * hh;
      if (BWD) {                                         // ATen clip_coordinates_set_grad
        gm |= (ix > 0.0f && ix < wm1) ? 1u << (2 * r) : 0u;
        gm |= (iy > 0.0f && iy < hm1) ? 2u << (2 * r) : 0u;
      }
      ix = fminf(fmaxf(ix, 0.0f), wm1);
      iy = fminf(fmaxf(iy, 0.0f), hm1);
      const int x0 = min(__float2int_rd(ix), W - 2), y0 = min(__float2int_rd(iy), H - 2);
      wxs[r] = ix - (float)x0;
      wys[r] = iy - (float)y0;
      xy[r] = (unsigned)(y0 * 65536 + x0);
    }
    __syncwarp();
    if (lx == 0) tc::mbar_arrive(flow_empty + 8 * fs);

    unsigned mask = 15u;
    if (!(tx * 32 + 32 <= W && ty * (8 * ROWS) + 8 * ROWS <= H)) {       // edge tile (CTA-uniform)
      mask = 0;
#pragma unroll
      for (int r = 0; r < ROWS; ++r) mask |= (px < W && py0 + 8 * r < H) ? 1u << r : 0u;
    }
    const unsigned pix0 = (unsigned)tb * 3u * plane + (unsigned)(py0 * W + px);   // only used where the mask bit is set
    unsigned pix[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) pix[r] = pix0 + (unsigned)(r * 8 * W);

    // ---- window of this tile
    const int ws = n % NWIN;
    w_win += wait_bar(win_full + 8 * ws, (uint32_t)(n / NWIN) & 1u);
    const int4 h = hdr[ws];
    const float* sb = reinterpret_cast<const float*>(sm + L::OFF_WIN + ws * L::WIN_BYTES);
    float v[ROWS][3][4];
    float g[ROWS][3];
    if (BWD) {
      const float* gsm = reinterpret_cast<const float*>(sm + L::OFF_GRAD + ws * GRAD_BYTES) + ly * 32 + lx;
#pragma unroll
      for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) g[r][ch] = gsm[ch * 1024 + r * 256];
    }
    if (h.z == 0) gather_staged<L::BOX_AW, L::BOX_AH>(sb, xy, h.x, h.y, v);
    else if (h.z == 1) gather_staged<L::BOX_BW, L::BOX_BH>(sb, xy, h.x, h.y, v);
    else {
      const float* const x0p = opaque(P.x + (size_t)tb * 3 * plane);
      const float* const x1p = opaque(x0p + plane);
      const float* const x2p = opaque(x1p + plane);
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        const unsigned a0 = (xy[r] >> 16) * (unsigned)W + (xy[r] & 0xffffu), a1 = a0 + (unsigned)W;
        v[r][0][0] = __ldg(x0p + a0); v[r][0][1] = __ldg(x0p + a0 + 1);
        v[r][0][2] = __ldg(x0p + a1); v[r][0][3] = __ldg(x0p + a1 + 1);
        v[r][1][0] = __ldg(x1p + a0); v[r][1][1] = __ldg(x1p + a0 + 1);
        v[r][1][2] = __ldg(x1p + a1); v[r][1][3] = __ldg(x1p + a1 + 1);
        v[r][2][0] = __ldg(x2p + a0); v[r][2][1] = __ldg(x2p + a0 + 1);
        v[r][2][2] = __ldg(x2p + a1); v[r][2][3] = __ldg(x2p + a1 + 1);
      }
    }

    if (!BWD) {
      // the blend consumes every tap, so the window can be handed back before the stores
      auto emit = [&](auto all) {
        float res[ROWS][3];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
          const float wx = wxs[r], wy = wys[r];
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            const float top = fmaf(wx, v[r][ch][1] - v[r][ch][0], v[r][ch][0]);
            const float bot = fmaf(wx, v[r][ch][3] - v[r][ch][2], v[r][ch][2]);
            res[r][ch] = fmaf(wy, bot - top, top);
          }
        }
        __syncwarp();
        if (lx == 0) tc::mbar_arrive(win_empty + 8 * ws);
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
          if (decltype(all)::value || (mask & (1u << r))) {
            __stcs(o0 + pix[r], res[r][0]);
            __stcs(o1 + pix[r], res[r][1]);
            __stcs(o2 + pix[r], res[r][2]);
          }
        }
      };
      if (mask == 15u) emit(std::true_type{}); else emit(std::false_type{});
    } else {
      const unsigned back = (unsigned)tb * plane;        // pix counts 3 planes per sample, dflow [B,2,H,W] has 2
      auto emit = [&](auto all) {
        float rx[ROWS], ry[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
          const float wx = wxs[r], wy = wys[r], ex = 1.0f - wx, ey = 1.0f - wy;
          float gx = 0.f, gy = 0.f;
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            gx = fmaf(g[r][ch], (v[r][ch][1] - v[r][ch][0]) * ey + (v[r][ch][3] - v[r][ch][2]) * wy, gx);
            gy = fmaf(g[r][ch], (v[r][ch][2] - v[r][ch][0]) * ex + (v[r][ch][3] - v[r][ch][1]) * wx, gy);
          }
          // ATen: grad_grid = gix * (W-1)/2 * clipmask ; norm_grid backward: / (W-1) * 2
          gx = (gm & (1u << (2 * r))) ? gx * (wm1 * 0.5f) : 0.0f;
          gy = (gm & (2u << (2 * r))) ? gy * (hm1 * 0.5f) : 0.0f;
          rx[r] = div_rn(gx, wm1, P.rcw) * 2.0f;
          ry[r] = div_rn(gy, hm1, P.rch) * 2.0f;
        }
        __syncwarp();
        if (lx == 0) tc::mbar_arrive(win_empty + 8 * ws);
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
          if (decltype(all)::value || (mask & (1u << r))) {
            const unsigned dp = pix[r] - back;
            __stcs(o0 + dp, rx[r]);
            __stcs(o1 + dp, ry[r]);
          }
        }
      };
      if (mask == 15u) emit(std::true_type{}); else emit(std::false_type{});
    }
    advance(tx, ty, tb);
  }
  if (P.prof != nullptr && threadIdx.x == 0) {
    long long* q = P.prof + (size_t)blockIdx.x * 8;
    q[4] = clock64() - t_start; q[5] = w_flow; q[6] = w_win; q[7] = my_tiles;
  }
}

// fp32 tiled tensor map without swizzle (rank 3 or 4); out-of-bounds elements read as zero
int make_map(CUtensorMap* m, const float* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
             const cuuint32_t* box) {
  tc::EncodeTiledFn enc = tc::get_encoder();
  if (enc == nullptr) {
    emip_set_error("cuTensorMapEncodeTiled not available from the driver");
    return EMIP_ENOSYS;
  }
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float*>(base), dims, strides_bytes, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    emip_set_error("flow_warp: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return EMIP_EINVAL;
  }
  return EMIP_OK;
}

int make_image_map(CUtensorMap* m, const float* x, int B, int H, int W, int bw, int bh) {
  cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B * 3};
  cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
  cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 3};
  return make_map(m, x, 3, dims, strides, box);
}

int make_flow_map(CUtensorMap* m, const float* flow, int B, int H, int W, long long fsb, long long fsc) {
  cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, 2, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)fsc * 4, (cuuint64_t)fsb * 4};
  cuuint32_t box[4] = {32, 32, 2, 1};
  return make_map(m, flow, 4, dims, strides, box);
}

long long* g_staged_prof = nullptr;   // diagnostic switch (tools/k3_roles.py), documented in include/emip_b200.h

}  // namespace

// Wait-cycle profile buffer: device pointer to [grid][8] int64 (producer: total, flow_empty, flow_full, win_empty;
// worker warp 0: total, flow_full, win_full, tiles) or null.  Diagnostic, not thread-safe.
extern "C" void emip_debug_flow_warp_staged_profile(long long* buf) { g_staged_prof = buf; }

// Stateless: launches the staged kernel or returns EMIP_ENOSYS when the shape / alignment is not covered.  The kernel-choice
// feedback lives with the CALLER: when dev_stats (device, 16 zeroed bytes, not shared between launches in flight) and
// host_stats (device-accessible pinned host memory, 8 bytes) are given, the last CTA publishes
// host_stats = {seq, per-mille of tiles that had to gather from global memory} -- the staged kernel is ~15 % faster than
// the direct-gather kernel on smooth flow fields and several times slower when most tiles fall back (noisy flow); both
// produce bit-identical results, so a caller may switch kernels on that number (emip_b200/warp.py does).
int flow_warp_staged_launch(bool bwd, const float* x, const float* flow, const float* dout, float* out, int B, int H, int W,
                            long long fsb, long long fsc, int ctas_per_sm, unsigned* dev_stats, unsigned* host_stats,
                            unsigned seq, cudaStream_t st) {
  // TMA: 16-byte aligned bases and pitches; tap coordinates are packed 16 bits per axis (dp2a operands are signed)
  auto misaligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) != 0; };
  if ((W & 3) != 0 || W >= 32768 || H >= 32768 || W < 2 || H < 2 || misaligned(x) || misaligned(flow) ||
      (bwd && misaligned(dout)) || fsc <= 0 || fsb <= 0 || (fsc & 3) != 0 || (fsb & 3) != 0 ||
      fsc < (long long)H * W || (B > 1 && fsb < 2 * fsc))
    return EMIP_ENOSYS;
  if ((long long)B * 3 * H * W >= 0x7fffffffLL) return EMIP_ENOSYS;       // 32-bit element offsets
  // Single-entry caches: the photometric loss warps the same image with the forward and the backward flow, and a
  // training loop reuses its buffers.  A map only depends on (pointer, shape, strides).
  struct Cache {
    const float* x; const float* flow; const float* dout;
    int B, H, W; long long fsb, fsc;
    StagedMaps maps;
  };
  static thread_local Cache caches[2] = {{nullptr, nullptr, nullptr, 0, 0, 0, 0, 0, {}}, {nullptr, nullptr, nullptr, 0, 0, 0, 0, 0, {}}};
  Cache& c = caches[bwd ? 1 : 0];                      // forward and backward use different window shapes
  if (c.B != B || c.H != H || c.W != W) { c.x = nullptr; c.flow = nullptr; c.dout = nullptr; c.B = B; c.H = H; c.W = W; }
  int rc;
  if (c.x != x) {
    c.x = nullptr;
    const int aw = bwd ? Layout<true>::BOX_AW : Layout<false>::BOX_AW, bw = bwd ? Layout<true>::BOX_BW : Layout<false>::BOX_BW;
    if ((rc = make_image_map(&c.maps.win[0], x, B, H, W, aw, Layout<true>::BOX_AH)) != EMIP_OK) return rc;
    if ((rc = make_image_map(&c.maps.win[1], x, B, H, W, bw, Layout<true>::BOX_BH)) != EMIP_OK) return rc;
    c.x = x;
  }
  if (c.flow != flow || c.fsb != fsb || c.fsc != fsc) {
    c.flow = nullptr;
    if ((rc = make_flow_map(&c.maps.flow, flow, B, H, W, fsb, fsc)) != EMIP_OK) return rc;
    c.flow = flow; c.fsb = fsb; c.fsc = fsc;
  }
  if (bwd && c.dout != dout) {
    c.dout = nullptr;
    if ((rc = make_image_map(&c.maps.grad, dout, B, H, W, 32, 32)) != EMIP_OK) return rc;
    c.dout = dout;
  }
  const int smem = bwd ? Layout<true>::BYTES : Layout<false>::BYTES;
  if (int rc__ = emip_func_max_smem(bwd ? (const void*)flow_warp_staged_kernel<true> : (const void*)flow_warp_staged_kernel<false>, smem)) return rc__;
  const int tiles_x = (W + 31) / 32, tiles_y = (H + 8 * ROWS - 1) / (8 * ROWS);
  const long long nblk = (long long)B * tiles_x * tiles_y;
  if (nblk >= 0x7fffffffLL) return EMIP_ENOSYS;
  const long long cap = (long long)emip_num_sms() * (ctas_per_sm == 1 ? 1 : 2);
  const int grid = (int)(nblk < cap ? nblk : cap);
  StagedParams P;
  P.x = x; P.out = out;
  P.H = H; P.W = W; P.tiles_x = tiles_x; P.tiles_y = tiles_y; P.n_tiles = (int)nblk;
  P.gx = grid % tiles_x; P.gy = (grid / tiles_x) % tiles_y; P.gb = grid / (tiles_x * tiles_y);
  P.rcw = 1.0f / (float)(W - 1); P.rch = 1.0f / (float)(H - 1);
  P.prof = g_staged_prof;
  P.dev_stats = nullptr; P.host_stats = nullptr; P.seq = 0;
  if (dev_stats != nullptr && host_stats != nullptr) {      // publishing costs ~3 us of kernel tail (system-scope write)
    P.seq = seq ? seq : 1u;
    P.dev_stats = dev_stats;
    P.host_stats = host_stats;
  }
  if (bwd) flow_warp_staged_kernel<true><<<grid, THREADS, smem, st>>>(c.maps, P);
  else flow_warp_staged_kernel<false><<<grid, THREADS, smem, st>>>(c.maps, P);
  return EMIP_OK;
}
