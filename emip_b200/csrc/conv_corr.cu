// f1: first layer of conv_corr on the NEVER-MATERIALISED cost volume.
//
// Replaces reference model/EMIP_short/model.py:59 (`nn.Conv2d(44*44, 968, 3, 1, 1)`, first layer of `conv_corr`)
// applied at model.py:96 to the `corr` tensor of matching.py:16-20, corr[b, j, y, x] = S[b, (y,x), j] =
// sum_c f0[b,c,(y,x)] f1[b,c,j] / sqrt(C).  Re-associating the two contractions,
//
//   out[b,o,y,x] = bias[o] + sum_{j,dy,dx} Wt[o,j,dy,dx] corr[b,j,y+dy-1,x+dx-1]            (zero padding in (y,x))
//                = bias[o] + sum_{dy,dx,c} f0[b,c,(y+dy-1,x+dx-1)] G[b][(o,dy,dx), c]
//   G[b][(o,dy,dx), c] = (1/sqrt(C)) sum_j Wt[o,j,dy,dx] f1[b,c,j]
//
// turns a 65.3 GFLOP/sample convolution over a 15 MB/sample input into two 4.3 GFLOP GEMMs per sample on the 1 MB
// feature maps, and removes the only consumer of the cost volume (the matching kernel can run flow-only).
//
// Both GEMMs run on the 5th-gen tensor cores in the same accuracy-preserving form as the matching kernel: operands
// are bf16 hi + lo pairs, D += A.hi B.hi + A.lo B.hi + A.hi B.lo with fp32 accumulation in TMEM (SS-mode tcgen05.mma,
// M = 128), operands staged by TMA into 128-byte-swizzled K-major tiles:
//   GEMM 1 (G):    A = prepared weight [(o,dy,dx)][j] (emip_conv_corr_prepare_weight: permuted + split once per
//                  weight version), B = f1 channel-major [c][j]; N = 128; the epilogue scales by 1/sqrt(C) and writes G
//                  as bf16 hi | lo, row (o,dy,dx) x column c -- which IS the K-major A operand of GEMM 2;
//   GEMM 2 (conv): A = G[b] [o][(dy,dx,c)], B = the token-major f0 read through a 4-D tensor map {c, x, y, b} with
//                  the box origin shifted by the tap (dx-1, dy-1): out-of-image pixels are zero-filled by the TMA
//                  unit, which is exactly the convolution's zero padding -- no im2col buffer; N = R image rows
//                  (R*W a multiple of 16, <= 256; 4 x 44 = 176 for the model's grid); epilogue adds the bias.
// The GEMM kernel itself is gemm_tc.cu (modes 0 and 1); this file is the C-ABI front end.
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/emip_b200.h"
#include "pair_common.cuh"
#include "match_tc.cuh"
#include "pair_bwd_tc.cuh"
#include "gemm_tc.cuh"
#include <math.h>

namespace {
using tc::TM;
constexpr int KCH = GEMM_TC_KCH;
constexpr int NMAX = 256;

// [O][N][3][3] fp32 -> hi | lo bf16 [2][O*9][ld], row (o, dy, dx), K = j contiguous
__global__ void __launch_bounds__(256)
prepare_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int O, int N,
                      long long ld) {
  const int o = blockIdx.y;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x) {
    const float* s = w + ((size_t)o * N + j) * 9;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float v = __ldg(s + t);
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      hi[((size_t)o * 9 + t) * ld + j] = h;
      lo[((size_t)o * 9 + t) * ld + j] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
}

long long weight_ld(int N) { return (N + 7) / 8 * 8; }      // 16-byte row pitch

// image rows per pixel tile: R*W a multiple of 16 and <= 256 (0: unsupported grid)
int rows_per_tile(int H, int W) {
  int best = 0;
  for (int r = 1; r * W <= NMAX && r <= H + 3; ++r)
    if ((r * W) % 16 == 0) best = r;
  return best;
}

struct Ws {
  void* f0_tok;        // bf16 [B][N][256]
  void* f1_chn;        // bf16 [B][256][ld]
  __nv_bfloat16* g;    // bf16 [2][B][O*9][128]
};
size_t g_bytes(int B, int O) { return emip_align_up((size_t)2 * B * O * 9 * 128 * 2, 1024); }
}  // namespace

extern "C" int emip_conv_corr_supported(int C, int H, int W) {
  return C == 128 && rows_per_tile(H, W) > 0 && H * W >= 16 ? 1 : 0;
}

extern "C" size_t emip_conv_corr_weight_bytes(int O, int N) {
  if (O <= 0 || N <= 0) return 0;
  return emip_align_up((size_t)2 * O * 9 * weight_ld(N) * 2, 1024);
}

extern "C" int emip_conv_corr_prepare_weight(const float* w, void* w_prep, int O, int N, void* stream) {
  EMIP_CHECK_ARG(w && w_prep && O > 0 && N > 0, "conv_corr_prepare_weight: bad arguments");
  EMIP_CHECK_ARG(reinterpret_cast<uintptr_t>(w_prep) % 1024 == 0, "conv_corr_prepare_weight: w_prep must be 1024-byte aligned");
  const long long ld = weight_ld(N);
  __nv_bfloat16* hi = static_cast<__nv_bfloat16*>(w_prep);
  __nv_bfloat16* lo = hi + (size_t)O * 9 * ld;
  cudaStream_t st = (cudaStream_t)stream;
  if (ld != N) EMIP_CUDA(cudaMemsetAsync(w_prep, 0, (size_t)2 * O * 9 * ld * 2, st));
  dim3 grid((N + 255) / 256, O);
  prepare_weight_kernel<<<grid, 256, 0, st>>>(w, hi, lo, O, N, ld);
  EMIP_CHECK_LAUNCH("conv_corr_prepare_weight");
  return EMIP_OK;
}

extern "C" size_t emip_conv_corr_workspace(int B, int C, int H, int W, int O) {
  if (B < 0 || C != 128 || H <= 0 || W <= 0 || O <= 0) return 0;
  return match_tc_split_bytes(B, H * W, C) + pair_bwd_tc_chn_bytes(B, H * W) + g_bytes(B, O);
}

extern "C" int emip_conv_corr_fwd(const float* f0, const float* f1, const void* w_prep, const float* bias, float* out,
                                  void* workspace, size_t ws_bytes, int B, int C, int H, int W, int O, void* stream) {
  return emip_conv_corr_fwd_ex(f0, f1, w_prep, bias, nullptr, nullptr, 0, EMIP_LAYOUT_CN, out, workspace, ws_bytes, B, C, H, W, O, stream);
}

// ep_scale != NULL: out = act(conv * ep_scale[o] + ep_shift[o]) -- the eval-mode BatchNorm2d + ReLU behind conv_corr[0]
// (model.py:60-61) folded into the epilogue, with the convolution's bias inside ep_shift (bias is then ignored)
extern "C" int emip_conv_corr_fwd_ex(const float* f0, const float* f1, const void* w_prep, const float* bias, const float* ep_scale,
                                     const float* ep_shift, int relu, int layout, float* out, void* workspace, size_t ws_bytes,
                                     int B, int C, int H, int W, int O, void* stream) {
  return emip_conv_corr_fwd_tokens(f0, f1, w_prep, bias, ep_scale, ep_shift, relu, layout, out, nullptr, workspace, ws_bytes, B, C, H, W, O,
                                   stream);
}

// tok_out != NULL (out may then be NULL): the result is written as the token-major bf16 hi | lo input operand of a following
// emip_conv3x3_fwd_tokens -- [B][H*W][2 * Cp] with Cp = O rounded up to 128, hi at [0, Cp), lo at [Cp, 2 Cp), padding channels
// zero -- instead of fp32 [B,O,H,W]: conv_corr[0] + BatchNorm + ReLU hand conv_corr[3] its operand without an fp32 round trip
extern "C" int emip_conv_corr_fwd_tokens(const float* f0, const float* f1, const void* w_prep, const float* bias, const float* ep_scale,
                                         const float* ep_shift, int relu, int layout, float* out, void* tok_out, void* workspace,
                                         size_t ws_bytes, int B, int C, int H, int W, int O, void* stream) {
  if (B == 0) return EMIP_OK;
  EMIP_CHECK_ARG(out != nullptr || tok_out != nullptr, "conv_corr_fwd: no output");
  EMIP_CHECK_ARG(tok_out == nullptr || reinterpret_cast<uintptr_t>(tok_out) % 1024 == 0, "conv_corr_fwd: tok_out must be 1024-byte aligned");
  EMIP_CHECK_ARG((ep_scale == nullptr) == (ep_shift == nullptr), "conv_corr_fwd: ep_scale and ep_shift come together");
  EMIP_CHECK_ARG(layout == EMIP_LAYOUT_NC || layout == EMIP_LAYOUT_CN, "conv_corr_fwd: layout must be 0 (token-major) or 1 (channel-major)");
  EMIP_CHECK_ARG(f0 && f1 && w_prep, "conv_corr_fwd: null pointer");
  EMIP_CHECK_ARG(B >= 0 && H > 0 && W > 0 && O > 0, "conv_corr_fwd: bad shape B=%d H=%d W=%d O=%d", B, H, W, O);
  if (!emip_conv_corr_supported(C, H, W)) {
    emip_set_error("conv_corr_fwd: unsupported shape C=%d H=%d W=%d (needs C=128 and R*W %% 16 == 0 for some R*W <= 256)", C, H, W);
    return EMIP_ENOSYS;
  }
  const int N = H * W;
  const size_t need = emip_conv_corr_workspace(B, C, H, W, O);
  if (workspace == nullptr || ws_bytes < need || reinterpret_cast<uintptr_t>(workspace) % 1024 != 0) {
    emip_set_error("conv_corr_fwd: workspace too small or not 1024-byte aligned (%zu < %zu bytes)", ws_bytes, need);
    return EMIP_ENOMEM;
  }
  cudaStream_t st = (cudaStream_t)stream;
  char* wsp = static_cast<char*>(workspace);
  Ws ws;
  ws.f0_tok = wsp;
  ws.f1_chn = wsp + match_tc_split_bytes(B, N, C);
  ws.g = reinterpret_cast<__nv_bfloat16*>(wsp + match_tc_split_bytes(B, N, C) + pair_bwd_tc_chn_bytes(B, N));
  int rc;
  if ((rc = match_tc_split(f0, nullptr, ws.f0_tok, B, N, C, layout, 0, st))) return rc;
  if ((rc = pair_bwd_tc_split_chn(f1, nullptr, ws.f1_chn, B, N, layout, st))) return rc;

  const int M1 = O * 9;
  const long long wld = weight_ld(N), cld = pair_bwd_tc_chn_ld(N);
  const __nv_bfloat16* w_hi = static_cast<const __nv_bfloat16*>(w_prep);
  const __nv_bfloat16* w_lo = w_hi + (size_t)M1 * wld;
  __nv_bfloat16* g_hi = ws.g;
  __nv_bfloat16* g_lo = ws.g + (size_t)B * M1 * 128;

  // ---- GEMM 1: G[b] = Wp f1[b]^T / sqrt(C)
  {
    CUtensorMap ma_hi, ma_lo, mb;
    const cuuint64_t adims[3] = {(cuuint64_t)N, (cuuint64_t)M1, 1}, astr[2] = {(cuuint64_t)wld * 2, (cuuint64_t)M1 * wld * 2};
    const cuuint32_t abox[3] = {KCH, TM, 1};
    if ((rc = gemm_tc_make_map(&ma_hi, w_hi, 3, adims, astr, abox))) return rc;
    if ((rc = gemm_tc_make_map(&ma_lo, w_lo, 3, adims, astr, abox))) return rc;
    const cuuint64_t bdims[3] = {(cuuint64_t)N, 256, (cuuint64_t)B}, bstr[2] = {(cuuint64_t)cld * 2, (cuuint64_t)cld * 2 * 256};
    const cuuint32_t bbox[3] = {KCH, 128, 1};
    if ((rc = gemm_tc_make_map(&mb, ws.f1_chn, 3, bdims, bstr, bbox))) return rc;
    GemmTcParams p = {};
    p.mode = 0; p.M = M1; p.n_mtiles = (M1 + TM - 1) / TM; p.n_ntiles = 1; p.n_tile = 128;
    p.kchunks = (N + KCH - 1) / KCH;
    p.scale = 1.0f / sqrtf((float)C);
    p.g_hi = g_hi; p.g_lo = g_lo;
    if ((rc = gemm_tc_launch(ma_hi, ma_lo, mb, p, B, st))) return rc;
  }
  // ---- GEMM 2: out[b] = conv3x3(f0[b]; G[b]) + bias
  {
    const int R = rows_per_tile(H, W);
    CUtensorMap ma_hi, ma_lo, mb;
    const cuuint64_t adims[3] = {9 * 128, (cuuint64_t)O, (cuuint64_t)B}, astr[2] = {9 * 128 * 2, (cuuint64_t)M1 * 128 * 2};
    const cuuint32_t abox[3] = {KCH, TM, 1};
    if ((rc = gemm_tc_make_map(&ma_hi, g_hi, 3, adims, astr, abox))) return rc;
    if ((rc = gemm_tc_make_map(&ma_lo, g_lo, 3, adims, astr, abox))) return rc;
    const cuuint64_t bdims[4] = {256, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t bstr[3] = {512, (cuuint64_t)W * 512, (cuuint64_t)N * 512};
    const cuuint32_t bbox[4] = {KCH, (cuuint32_t)W, (cuuint32_t)R, 1};
    if ((rc = gemm_tc_make_map(&mb, ws.f0_tok, 4, bdims, bstr, bbox))) return rc;
    GemmTcParams p = {};
    p.mode = 1; p.M = O; p.n_mtiles = (O + TM - 1) / TM; p.n_ntiles = (H + R - 1) / R; p.n_tile = R * W;
    p.kchunks = 18; p.cpt = 2; p.lo_off = 128;
    p.W = W; p.H = H; p.R = R;
    p.bias = ep_scale ? ep_shift : bias; p.ep_scale = ep_scale; p.ep_relu = relu; p.out = out;
    if (tok_out != nullptr) {
      const int Cp = (O + TM - 1) / TM * TM;
      p.tok_hi = static_cast<__nv_bfloat16*>(tok_out); p.tok_ld = 2 * Cp; p.tok_lo_off = Cp;
    }
    if ((rc = gemm_tc_launch(ma_hi, ma_lo, mb, p, B, st))) return rc;
  }
  return EMIP_OK;
}

// Backward of emip_conv_corr_fwd on the tensor cores (five mode-2 GEMMs, gemm_tc.cu): df0, df1 [B,C,H,W],
// dweight [O,H*W,3,3], dbias [O] (NULL: not wanted) from dout [B,O,H,W].  weight = the fp32 parameter, w_prep = its
// prepared copy (emip_conv_corr_prepare_weight).
extern "C" size_t emip_conv_corr_bwd_workspace(int B, int C, int H, int W, int O) {
  if (B < 0 || C != 128 || H <= 0 || W <= 0 || O <= 0) return 0;
  return conv_corr_bwd_scratch_bytes(B, O, H * W);
}

extern "C" int emip_conv_corr_bwd(const float* f0, const float* f1, const float* weight, const void* w_prep, const float* dout,
                                  float* df0, float* df1, float* dweight, float* dbias, void* workspace, size_t ws_bytes, int B,
                                  int C, int H, int W, int O, void* stream) {
  if (B == 0) return EMIP_OK;
  EMIP_CHECK_ARG(f0 && f1 && weight && w_prep && dout && df0 && df1 && dweight && workspace, "conv_corr_bwd: null pointer");
  if (!emip_conv_corr_supported(C, H, W)) {
    emip_set_error("conv_corr_bwd: unsupported shape C=%d H=%d W=%d", C, H, W);
    return EMIP_ENOSYS;
  }
  return conv_corr_bwd_tc(f0, f1, weight, w_prep, weight_ld(H * W), dout, df0, df1, dweight, dbias, workspace, ws_bytes, B, H, W, O,
                          (cudaStream_t)stream);
}
