// The 3x3 convolutions on the chained path (VERDICT r1 row g) on the tensor cores, and the channel-major -> token-major
// hand-over between the prompt fusion and the FeatureTransformer.
//
//   * emip_conv3x3_fwd: out[b,o,y,x] = act(scale[o] * sum_{c,dy,dx} w[o,c,dy,dx] in[b,c,y+dy-1,x+dx-1] + shift[o]), zero
//     padding, stride 1 -- reference model/EMIP_short/motion/gmflow/gmflow.py:43-44 (upsampler[0]: Conv2d(2 + 128, 256, 3, 1, 1)
//     + ReLU on cat(flow, feature), :62-64) and model/EMIP_short/model.py:62 (conv_corr[3]: Conv2d(968, 128, 3, 1, 1)).
//     The convolution is the mode-1 GEMM of gemm_tc.cu: A = the prepared weight [o][(tap, c)] (bf16 hi | lo, K-major), B =
//     the token-major input read through a 4-D tensor map {c, x, y, b} whose box origin is shifted by the tap -- the TMA
//     unit's zero fill is the zero padding, there is no im2col buffer; 3-term split-bf16 UMMAs, fp32 accumulation in TMEM.
//     The input is the channel concatenation of up to two sources, each channel-major [B][C][H*W] or token-major
//     [B][H*W][C] (the transformer's output): one split pass writes the bf16 hi | lo token rows.
//   * emip_tokens_from_cn: out[b][n][c] = x[b][c][n] + pos[c][n] -- gmflow.py:114 (feature_add_position, utils.py:66-86)
//     + transformer.py:439-440 (flatten / permute) in one pass.
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/emip_b200.h"
#include "pair_common.cuh"
#include "gemm_tc.cuh"
#include <math.h>

namespace {
using tc::TM;
constexpr int KCH = GEMM_TC_KCH;
constexpr int NMAX = 256;

int cpad(int c) { return (c + KCH - 1) / KCH * KCH; }

int rows_per_tile(int H, int W) {
  int best = 0;
  for (int r = 1; r * W <= NMAX && r <= H + 3; ++r)
    if ((r * W) % 16 == 0) best = r;
  return best;
}

// [O][Cin][3][3] fp32 -> hi | lo bf16 [2][O][9 * Cp], K index = tap * Cp + c (zero padded channels)
__global__ void __launch_bounds__(256)
conv3x3_prepare_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int O, int Cin,
                       int Cp) {
  const int o = blockIdx.y;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < Cp; c += gridDim.x * blockDim.x) {
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float v = c < Cin ? __ldg(w + ((size_t)o * Cin + c) * 9 + t) : 0.f;
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      hi[((size_t)o * 9 + t) * Cp + c] = h;
      lo[((size_t)o * 9 + t) * Cp + c] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
}

struct SplitSrc {
  const float* x[2];
  int C[2];
  int layout[2];
};

// cat(x0, x1) over channels -> token-major bf16 [B][N][2 * Cp] (hi | lo): one CTA = 32 tokens x 64 channels through a
// padded shared-memory tile (coalesced along n for channel-major sources, along c for token-major ones)
__global__ void __launch_bounds__(256)
conv_split_tok_kernel(const SplitSrc s, __nv_bfloat16* __restrict__ dst, int N, int Cp) {
  __shared__ float t[64][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 64, n0 = blockIdx.x * 32;
  const int Ct = s.C[0] + s.C[1];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;           // 8 warps
  // channel-major reads: lane = token, warp walks channels; token-major reads: lane = channel pair, warp walks tokens
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int cl = ty * 8 + i, c = c0 + cl;
    float v = 0.f;
    if (c < Ct && n0 + tx < N) {
      const int si = c < s.C[0] ? 0 : 1, cc = c - (si ? s.C[0] : 0);
      if (s.layout[si] == EMIP_LAYOUT_CN) v = __ldg(s.x[si] + ((size_t)b * s.C[si] + cc) * N + n0 + tx);
    }
    t[cl][tx] = v;
  }
  __syncthreads();
  // token-major sources: coalesced along c (overwrites the zeros written above for their channels)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int nl = ty * 4 + i, n = n0 + nl;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int cl = tx + 32 * h, c = c0 + cl;
      if (c < Ct && n < N) {
        const int si = c < s.C[0] ? 0 : 1, cc = c - (si ? s.C[0] : 0);
        if (s.layout[si] == EMIP_LAYOUT_NC) t[cl][nl] = __ldg(s.x[si] + ((size_t)b * N + n) * s.C[si] + cc);
      }
    }
  }
  __syncthreads();
  // write: thread = (token, channel pair): a warp writes 64 B hi + 64 B lo runs of one token row
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int nl = ty * 4 + i, n = n0 + nl;
    if (n < N) {
      const float v0 = t[2 * tx][nl], v1 = t[2 * tx + 1][nl];
      const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
      __nv_bfloat16* d = dst + ((size_t)b * N + n) * 2 * Cp + c0 + 2 * tx;
      *reinterpret_cast<__nv_bfloat162*>(d) = __halves2bfloat162(h0, h1);
      *reinterpret_cast<__nv_bfloat162*>(d + Cp) =
          __halves2bfloat162(__float2bfloat16_rn(v0 - __bfloat162float(h0)), __float2bfloat16_rn(v1 - __bfloat162float(h1)));
    }
  }
}

// out[b][n][c] = x[b][c][n] (+ pos[c][n]): 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256)
tokens_from_cn_kernel(const float* __restrict__ x, const float* __restrict__ pos, float* __restrict__ out, int C, int N) {
  __shared__ float t[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty * 4 + i, n = n0 + tx;
    if (c < C && n < N) t[ty * 4 + i][tx] = __ldg(x + ((size_t)b * C + c) * N + n) + (pos ? __ldg(pos + (size_t)c * N + n) : 0.f);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i, c = c0 + tx;
    if (c < C && n < N) out[((size_t)b * N + n) * C + c] = t[tx][ty * 4 + i];
  }
}
}  // namespace

extern "C" int emip_conv3x3_supported(int Cin, int H, int W) {
  return Cin >= 1 && Cin <= 4096 && rows_per_tile(H, W) > 0 && H * W >= 16 ? 1 : 0;
}

extern "C" size_t emip_conv3x3_weight_bytes(int O, int Cin) {
  if (O <= 0 || Cin <= 0) return 0;
  return emip_align_up((size_t)2 * O * 9 * cpad(Cin) * 2, 1024);
}

extern "C" int emip_conv3x3_prepare_weight(const float* w, void* w_prep, int O, int Cin, void* stream) {
  EMIP_CHECK_ARG(w && w_prep && O > 0 && Cin > 0, "conv3x3_prepare_weight: bad arguments");
  EMIP_CHECK_ARG(reinterpret_cast<uintptr_t>(w_prep) % 1024 == 0, "conv3x3_prepare_weight: w_prep must be 1024-byte aligned");
  const int Cp = cpad(Cin);
  __nv_bfloat16* hi = static_cast<__nv_bfloat16*>(w_prep);
  __nv_bfloat16* lo = hi + (size_t)O * 9 * Cp;
  conv3x3_prepare_kernel<<<dim3((Cp + 255) / 256, O), 256, 0, (cudaStream_t)stream>>>(w, hi, lo, O, Cin, Cp);
  EMIP_CHECK_LAUNCH("conv3x3_prepare_weight");
  return EMIP_OK;
}

extern "C" size_t emip_conv3x3_workspace(int B, int Cin, int H, int W) {
  if (B < 0 || Cin <= 0 || H <= 0 || W <= 0) return 0;
  return emip_align_up((size_t)B * H * W * 2 * cpad(Cin) * 2, 1024);
}

extern "C" int emip_conv3x3_fwd(const float* x0, int C0, int layout0, const float* x1, int C1, int layout1, const void* w_prep,
                                const float* scale, const float* shift, int relu, float* out, void* workspace, size_t ws_bytes,
                                int B, int H, int W, int O, void* stream) {
  if (B == 0) return EMIP_OK;
  EMIP_CHECK_ARG(x0 && C0 > 0 && w_prep && out && workspace, "conv3x3_fwd: null pointer");
  EMIP_CHECK_ARG((x1 == nullptr) == (C1 == 0) && C1 >= 0, "conv3x3_fwd: x1 and C1 come together");
  EMIP_CHECK_ARG(B > 0 && H > 0 && W > 0 && O > 0, "conv3x3_fwd: bad shape B=%d H=%d W=%d O=%d", B, H, W, O);
  EMIP_CHECK_ARG((layout0 == EMIP_LAYOUT_NC || layout0 == EMIP_LAYOUT_CN) && (layout1 == EMIP_LAYOUT_NC || layout1 == EMIP_LAYOUT_CN),
                 "conv3x3_fwd: layout must be 0 (token-major) or 1 (channel-major)");
  const int Cin = C0 + C1, N = H * W;
  if (!emip_conv3x3_supported(Cin, H, W)) {
    emip_set_error("conv3x3_fwd: unsupported shape Cin=%d H=%d W=%d (needs R*W %% 16 == 0 for some R*W <= 256)", Cin, H, W);
    return EMIP_ENOSYS;
  }
  if (ws_bytes < emip_conv3x3_workspace(B, Cin, H, W) || reinterpret_cast<uintptr_t>(workspace) % 1024 != 0) {
    emip_set_error("conv3x3_fwd: workspace too small or not 1024-byte aligned");
    return EMIP_ENOMEM;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int Cp = cpad(Cin), R = rows_per_tile(H, W);
  __nv_bfloat16* tok = static_cast<__nv_bfloat16*>(workspace);
  SplitSrc s;
  s.x[0] = x0; s.C[0] = C0; s.layout[0] = layout0;
  s.x[1] = x1; s.C[1] = C1; s.layout[1] = layout1;
  conv_split_tok_kernel<<<dim3((N + 31) / 32, Cp / 64, B), 256, 0, st>>>(s, tok, N, Cp);
  EMIP_CHECK_LAUNCH("conv3x3_fwd (split)");
  const __nv_bfloat16* w_hi = static_cast<const __nv_bfloat16*>(w_prep);
  const __nv_bfloat16* w_lo = w_hi + (size_t)O * 9 * Cp;
  int rc;
  CUtensorMap ma_hi, ma_lo, mb;
  const cuuint64_t adims[3] = {(cuuint64_t)9 * Cp, (cuuint64_t)O, 1}, astr[2] = {(cuuint64_t)9 * Cp * 2, (cuuint64_t)O * 9 * Cp * 2};
  const cuuint32_t abox[3] = {KCH, TM, 1};
  if ((rc = gemm_tc_make_map(&ma_hi, w_hi, 3, adims, astr, abox))) return rc;
  if ((rc = gemm_tc_make_map(&ma_lo, w_lo, 3, adims, astr, abox))) return rc;
  const cuuint64_t bdims[4] = {(cuuint64_t)2 * Cp, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t bstr[3] = {(cuuint64_t)2 * Cp * 2, (cuuint64_t)W * 2 * Cp * 2, (cuuint64_t)N * 2 * Cp * 2};
  const cuuint32_t bbox[4] = {KCH, (cuuint32_t)W, (cuuint32_t)R, 1};
  if ((rc = gemm_tc_make_map(&mb, tok, 4, bdims, bstr, bbox))) return rc;
  GemmTcParams p = {};
  p.mode = 1; p.M = O; p.n_mtiles = (O + TM - 1) / TM; p.n_ntiles = (H + R - 1) / R; p.n_tile = R * W;
  p.cpt = Cp / KCH; p.kchunks = 9 * p.cpt; p.lo_off = Cp; p.a_shared = 1;
  p.W = W; p.H = H; p.R = R;
  p.bias = shift; p.ep_scale = scale; p.ep_relu = relu ? 1 : 0; p.out = out;
  return gemm_tc_launch(ma_hi, ma_lo, mb, p, B, st);
}

// As emip_conv3x3_fwd for an input that already is the token-major bf16 hi | lo operand ([B][H*W][2 * Cp], Cp = Cp_in: what
// emip_conv_corr_fwd_tokens writes): no split pass, no workspace.  w_prep = emip_conv3x3_prepare_weight(w, O, Cin) with
// cpad(Cin) == Cp_in.
extern "C" int emip_conv3x3_fwd_tokens(const void* tok, int Cin, int Cp_in, const void* w_prep, const float* scale, const float* shift,
                                       int relu, float* out, int B, int H, int W, int O, void* stream) {
  if (B == 0) return EMIP_OK;
  EMIP_CHECK_ARG(tok && w_prep && out && Cin > 0, "conv3x3_fwd_tokens: null pointer");
  EMIP_CHECK_ARG(Cp_in % KCH == 0 && Cp_in >= Cin && cpad(Cin) == Cp_in, "conv3x3_fwd_tokens: Cp_in must be Cin rounded up to %d", KCH);
  EMIP_CHECK_ARG(reinterpret_cast<uintptr_t>(tok) % 128 == 0, "conv3x3_fwd_tokens: tok must be 128-byte aligned");
  if (!emip_conv3x3_supported(Cin, H, W)) { emip_set_error("conv3x3_fwd_tokens: unsupported shape Cin=%d H=%d W=%d", Cin, H, W); return EMIP_ENOSYS; }
  cudaStream_t st = (cudaStream_t)stream;
  const int Cp = Cp_in, R = rows_per_tile(H, W), N = H * W;
  const __nv_bfloat16* w_hi = static_cast<const __nv_bfloat16*>(w_prep);
  const __nv_bfloat16* w_lo = w_hi + (size_t)O * 9 * Cp;
  int rc;
  CUtensorMap ma_hi, ma_lo, mb;
  const cuuint64_t adims[3] = {(cuuint64_t)9 * Cp, (cuuint64_t)O, 1}, astr[2] = {(cuuint64_t)9 * Cp * 2, (cuuint64_t)O * 9 * Cp * 2};
  const cuuint32_t abox[3] = {KCH, TM, 1};
  if ((rc = gemm_tc_make_map(&ma_hi, w_hi, 3, adims, astr, abox))) return rc;
  if ((rc = gemm_tc_make_map(&ma_lo, w_lo, 3, adims, astr, abox))) return rc;
  const cuuint64_t bdims[4] = {(cuuint64_t)2 * Cp, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t bstr[3] = {(cuuint64_t)2 * Cp * 2, (cuuint64_t)W * 2 * Cp * 2, (cuuint64_t)N * 2 * Cp * 2};
  const cuuint32_t bbox[4] = {KCH, (cuuint32_t)W, (cuuint32_t)R, 1};
  if ((rc = gemm_tc_make_map(&mb, tok, 4, bdims, bstr, bbox))) return rc;
  GemmTcParams p = {};
  p.mode = 1; p.M = O; p.n_mtiles = (O + TM - 1) / TM; p.n_ntiles = (H + R - 1) / R; p.n_tile = R * W;
  p.cpt = Cp / KCH; p.kchunks = 9 * p.cpt; p.lo_off = Cp; p.a_shared = 1;
  p.W = W; p.H = H; p.R = R;
  p.bias = shift; p.ep_scale = scale; p.ep_relu = relu ? 1 : 0; p.out = out;
  return gemm_tc_launch(ma_hi, ma_lo, mb, p, B, st);
}

extern "C" int emip_tokens_from_cn(const float* x, const float* pos, float* out, int B, int C, int N, void* stream) {
  if (B == 0) return EMIP_OK;
  EMIP_CHECK_ARG(x && out && B > 0 && C > 0 && N > 0, "tokens_from_cn: bad arguments");
  tokens_from_cn_kernel<<<dim3((N + 31) / 32, (C + 31) / 32, B), 256, 0, (cudaStream_t)stream>>>(x, pos, out, C, N);
  EMIP_CHECK_LAUNCH("tokens_from_cn");
  return EMIP_OK;
}
