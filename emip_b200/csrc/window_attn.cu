// f2: split-window attention of the GMFlow FeatureTransformer as ONE C-ABI call per layer,
// reference model/EMIP_short/motion/gmflow/transformer.py:46-105 (single_head_split_window_attention; mask :19-43).
//
// The reference splits the h x w token map into num_splits^2 windows (rolled by half a window when shifted, with a
// 0 / -100 additive mask), which makes 2 gathers + 1 scatter per tensor around the attention.  The mask only separates
// rectangular blocks of tokens (un-rolled coordinates: cuts at `shift` and `size - win + shift`), and e^{-100} relative
// weight is below fp32 resolution, so every layer is plain attention inside (num_splits + [shifted])^2 rectangular
// blocks.  Here the block geometry lives in the kernels: the operand-split pass reads q, k, v straight from the
// [B][h*w][128] tensors (window gather = address arithmetic) and the attention epilogue writes each row to its pixel
// of the [B][h*w][128] output (attn_tc.cu, AttnWinMap).  Blocks with the same token count share a launch: 2 launches
// for an unshifted layer, 6 for a shifted one (484-, 242- and 121-token blocks at 44 x 44) -- no roll, no mask tensor,
// no gather / scatter copies.
#include "common.cuh"
#include "../../include/emip_b200.h"
#include "pair_common.cuh"
#include "match_tc.cuh"
#include "pair_bwd_tc.cuh"
#include "attn_tc.cuh"
#include "mlp_fused.cuh"
#include "attn_bwd_tc.cuh"
#include "gemm_tc.cuh"
#include <math.h>

namespace {
constexpr int KC = 128;
constexpr int TOK = 64;                  // tokens per CTA of the split pass
constexpr int MAXBLK = 16;

struct Group {                           // blocks with the same number of tokens share a launch
  int n, nblk;
  int r0[MAXBLK], c0[MAXBLK], bw[MAXBLK];
};

// Ranges (un-rolled coordinates) of one axis inside which the reference lets tokens attend to each other.
int axis_groups(int size, int splits, int shift, int* a0, int* b0) {
  const int win = size / splits;
  int n = 0;
  if (shift == 0) {
    for (int i = 0; i < splits; ++i) { a0[n] = i * win; b0[n] = (i + 1) * win; ++n; }
    return n;
  }
  // rolled coordinate r = (orig - shift) mod size; windows [i*win, (i+1)*win) in rolled coordinates; the last window is
  // cut at size - shift by the mask labels of generate_shift_window_attn_mask (transformer.py:25-30)
  for (int i = 0; i < splits + 1; ++i) {
    const int a = i < splits - 1 ? i * win : (i == splits - 1 ? size - win : size - shift);
    const int b = i < splits - 1 ? (i + 1) * win : (i == splits - 1 ? size - shift : size);
    a0[n] = (a + shift) % size;
    b0[n] = (b + shift - 1) % size + 1;
    ++n;
  }
  return n;
}

// Blocks grouped by shape; returns the number of groups or -1 when a group would exceed MAXBLK blocks.
int make_groups(int h, int w, int splits, int with_shift, Group* g, int max_groups) {
  int ra[8], rb[8], ca[8], cb[8];
  if (splits < 1 || splits > 6 || h % splits || w % splits) return -1;
  const int sh = with_shift ? (h / splits) / 2 : 0, sw = with_shift ? (w / splits) / 2 : 0;
  const int nr = axis_groups(h, splits, sh, ra, rb), nc = axis_groups(w, splits, sw, ca, cb);
  int ng = 0;
  for (int i = 0; i < nr; ++i)
    for (int j = 0; j < nc; ++j) {
      const int bh = rb[i] - ra[i], bw = cb[j] - ca[j];
      if (bh <= 0 || bw <= 0) return -1;
      int k = 0;
      while (k < ng && g[k].n != bh * bw) ++k;
      if (k == ng) {
        if (ng == max_groups) return -1;
        g[ng].n = bh * bw; g[ng].nblk = 0;
        ++ng;
      }
      if (g[k].nblk == MAXBLK) return -1;
      g[k].r0[g[k].nblk] = ra[i];
      g[k].c0[g[k].nblk] = ca[j];
      g[k].bw[g[k].nblk] = bw;
      ++g[k].nblk;
    }
  return ng;
}

struct SplitParams {
  const float* src[4];       // q, k, v (, dout): [B][h*w][128]
  __nv_bfloat16* dst[4];     // [nprob][n][256] token-major hi|lo
  int B, h, w, n;
  int kv_shift;              // k, v (tensors 1, 2) of image i are read from image (i + kv_shift) % B
  int r0[MAXBLK], c0[MAXBLK], bw[MAXBLK];
};

// One launch splits all three operands of a block group: blockIdx = (64-token slab, problem, tensor).  One thread =
// one token x 4 channels per step: a warp reads one 512-byte token row and writes its 256-byte hi and lo halves.
__global__ void __launch_bounds__(256)
win_split_kernel(const __grid_constant__ SplitParams sp) {
  const int z = blockIdx.z, prob = blockIdx.y, t0 = blockIdx.x * TOK;
  const int blk = prob / sp.B, img = prob - blk * sp.B;
  const int simg = (z == 1 || z == 2) ? (img + sp.kv_shift) % sp.B : img;
  const float* src = sp.src[z] + (size_t)simg * sp.h * sp.w * KC;
  __nv_bfloat16* dst = sp.dst[z] + (size_t)prob * sp.n * 256;
  const int r0 = sp.r0[blk], c0 = sp.c0[blk], bw = sp.bw[blk];
  const int c4 = threadIdx.x & 31, tw = threadIdx.x >> 5;           // 8 warps x 8 tokens each
  float4 v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int t = t0 + tw * 8 + j;
    if (t < sp.n) {
      const int ty = t / bw, tx = t - ty * bw;
      v[j] = __ldg(reinterpret_cast<const float4*>(src + ((size_t)(r0 + ty) * sp.w + c0 + tx) * KC) + c4);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int t = t0 + tw * 8 + j;
    if (t < sp.n) {
      const __nv_bfloat16 h0 = __float2bfloat16_rn(v[j].x), h1 = __float2bfloat16_rn(v[j].y), h2 = __float2bfloat16_rn(v[j].z),
                          h3 = __float2bfloat16_rn(v[j].w);
      __nv_bfloat162 hi[2] = {__halves2bfloat162(h0, h1), __halves2bfloat162(h2, h3)};
      __nv_bfloat162 lo[2] = {__halves2bfloat162(__float2bfloat16_rn(v[j].x - __bfloat162float(h0)),
                                                 __float2bfloat16_rn(v[j].y - __bfloat162float(h1))),
                              __halves2bfloat162(__float2bfloat16_rn(v[j].z - __bfloat162float(h2)),
                                                 __float2bfloat16_rn(v[j].w - __bfloat162float(h3)))};
      *reinterpret_cast<uint2*>(dst + (size_t)t * 256 + 4 * c4) = *reinterpret_cast<uint2*>(hi);
      *reinterpret_cast<uint2*>(dst + (size_t)t * 256 + 128 + 4 * c4) = *reinterpret_cast<uint2*>(lo);
    }
  }
}

size_t group_bytes(int B, const Group& g) {
  return 3 * match_tc_split_bytes(g.nblk * B, g.n, KC);
}
size_t alf(size_t n_floats) { return emip_align_up(n_floats * sizeof(float), 1024); }
size_t group_bytes_bwd(int B, const Group& g);
size_t ws_bytes_groups(int B, const Group* g, int ng) {
  size_t need = 0;
  for (int i = 0; i < ng; ++i) {
    const size_t b = group_bytes_bwd(B, g[i]);
    need = b > need ? b : need;
  }
  return need;
}
// backward: four operand splits
size_t group_bytes_bwd(int B, const Group& g) { return 4 * match_tc_split_bytes(g.nblk * B, g.n, KC); }
}  // namespace

extern "C" size_t emip_window_attention_tc_workspace(int B, int h, int w, int C, int num_splits, int with_shift) {
  Group g[16];
  if (B < 0 || C != KC || h <= 0 || w <= 0) return 0;
  const int ng = make_groups(h, w, num_splits, with_shift, g, 16);
  if (ng < 0) return 0;
  size_t need = 0;
  for (int i = 0; i < ng; ++i) {
    const size_t b = group_bytes(B, g[i]);
    need = b > need ? b : need;
  }
  return need;
}

extern "C" int emip_window_attention_fwd_tc(const float* q, const float* k, const float* v, float* out, float* lse,
                                            void* workspace, size_t ws_bytes, int B, int h, int w, int C, int num_splits,
                                            int with_shift, void* stream) {
  return emip_window_attention_fwd_tc_ex(q, k, v, out, lse, workspace, ws_bytes, B, h, w, C, num_splits, with_shift, 0, stream);
}

// EMIP_WINATTN_KV_SWAP_HALVES: image i attends to the keys / values of image (i + B/2) % B -- the cross-attention layers of
// the FeatureTransformer, whose target is the source with the two batch halves swapped (transformer.py:462, :473): the
// swapped copy `concat1` is never made
extern "C" int emip_window_attention_fwd_tc_ex(const float* q, const float* k, const float* v, float* out, float* lse,
                                               void* workspace, size_t ws_bytes, int B, int h, int w, int C, int num_splits,
                                               int with_shift, int flags, void* stream) {
  if (B == 0) return EMIP_OK;
  EMIP_CHECK_ARG((flags & ~EMIP_WINATTN_KV_SWAP_HALVES) == 0, "window_attention_fwd_tc: unknown flag");
  EMIP_CHECK_ARG(!(flags & EMIP_WINATTN_KV_SWAP_HALVES) || B % 2 == 0, "window_attention_fwd_tc: KV_SWAP_HALVES needs an even batch");
  EMIP_CHECK_ARG(q && k && v && out && workspace, "window_attention_fwd_tc: null pointer");
  EMIP_CHECK_ARG(B > 0 && h > 0 && w > 0, "window_attention_fwd_tc: bad shape B=%d h=%d w=%d", B, h, w);
  if (C != KC) {
    emip_set_error("window_attention_fwd_tc: C=%d unsupported (kernels are built for the model's C=128)", C);
    return EMIP_ENOSYS;
  }
  Group g[16];
  const int ng = make_groups(h, w, num_splits, with_shift, g, 16);
  if (ng < 0) {
    emip_set_error("window_attention_fwd_tc: unsupported window geometry h=%d w=%d num_splits=%d", h, w, num_splits);
    return EMIP_ENOSYS;
  }
  if (ws_bytes < emip_window_attention_tc_workspace(B, h, w, C, num_splits, with_shift) ||
      reinterpret_cast<uintptr_t>(workspace) % 1024 != 0) {
    emip_set_error("window_attention_fwd_tc: workspace too small or not 1024-byte aligned");
    return EMIP_ENOMEM;
  }
  cudaStream_t st = (cudaStream_t)stream;
  for (int i = 0; i < ng; ++i) {
    const int n = g[i].n, nprob = g[i].nblk * B;
    char* base = static_cast<char*>(workspace);
    const size_t tok_bytes = match_tc_split_bytes(nprob, n, KC);
    SplitParams sp;
    sp.src[0] = q; sp.src[1] = k; sp.src[2] = v;
    sp.dst[0] = reinterpret_cast<__nv_bfloat16*>(base);
    sp.dst[1] = reinterpret_cast<__nv_bfloat16*>(base + tok_bytes);
    sp.dst[2] = reinterpret_cast<__nv_bfloat16*>(base + 2 * tok_bytes);
    sp.B = B; sp.h = h; sp.w = w; sp.n = n;
    sp.kv_shift = (flags & EMIP_WINATTN_KV_SWAP_HALVES) ? B / 2 : 0;
    AttnTcArgs a = {};
    a.win.enabled = 1; a.win.B = B; a.win.h = h; a.win.w = w;
    for (int j = 0; j < MAXBLK; ++j) {
      sp.r0[j] = a.win.r0[j] = j < g[i].nblk ? g[i].r0[j] : 0;
      sp.c0[j] = a.win.c0[j] = j < g[i].nblk ? g[i].c0[j] : 0;
      sp.bw[j] = a.win.bw[j] = j < g[i].nblk ? g[i].bw[j] : 1;
    }
    win_split_kernel<<<dim3((n + TOK - 1) / TOK, nprob, 3), 256, 0, st>>>(sp);
    EMIP_CHECK_LAUNCH("window_attention (split)");
    a.q_split = sp.dst[0]; a.k_split = sp.dst[1]; a.v_split = sp.dst[2];
    a.out = out; a.out_stride_b = 0; a.lse = lse;
    a.nb = nprob; a.nq = n; a.nk = n; a.out_layout = EMIP_LAYOUT_NC;
    a.sqrt_c = sqrtf((float)KC);
    a.ksplit = 1;
    int rc;
    if ((rc = attn_tc_fwd(a, st))) return rc;
  }
  return EMIP_OK;
}

extern "C" size_t emip_window_attention_bwd_tc_workspace(int B, int h, int w, int C, int num_splits, int with_shift) {
  Group g[16];
  if (B < 0 || C != KC || h <= 0 || w <= 0) return 0;
  const int ng = make_groups(h, w, num_splits, with_shift, g, 16);
  if (ng < 0) return 0;
  return ws_bytes_groups(B, g, ng) + alf((size_t)B * h * w);      // + D in pixel order
}

// Backward of emip_window_attention_fwd_tc: dq, dk, dv [B][h*w][C] from q, k, v, the forward's out and lse ([B][h*w], as
// the forward wrote it) and dout.  D = rowsum(dO o O) is formed once in pixel order; per block group one split launch
// (window gather of q, k, v, dout) and the three gradient launches, whose statistics loads and epilogues use the same
// pixel map.
extern "C" int emip_window_attention_bwd_tc(const float* q, const float* k, const float* v, const float* out, const float* lse,
                                            const float* dout, float* dq, float* dk, float* dv, void* workspace,
                                            size_t ws_bytes, int B, int h, int w, int C, int num_splits, int with_shift,
                                            void* stream) {
  if (B == 0) return EMIP_OK;
  EMIP_CHECK_ARG(q && k && v && out && lse && dout && dq && dk && dv && workspace, "window_attention_bwd_tc: null pointer");
  EMIP_CHECK_ARG(B > 0 && h > 0 && w > 0, "window_attention_bwd_tc: bad shape B=%d h=%d w=%d", B, h, w);
  if (C != KC) {
    emip_set_error("window_attention_bwd_tc: C=%d unsupported (kernels are built for the model's C=128)", C);
    return EMIP_ENOSYS;
  }
  Group g[16];
  const int ng = make_groups(h, w, num_splits, with_shift, g, 16);
  if (ng < 0) {
    emip_set_error("window_attention_bwd_tc: unsupported window geometry h=%d w=%d num_splits=%d", h, w, num_splits);
    return EMIP_ENOSYS;
  }
  if (ws_bytes < emip_window_attention_bwd_tc_workspace(B, h, w, C, num_splits, with_shift) ||
      reinterpret_cast<uintptr_t>(workspace) % 1024 != 0) {
    emip_set_error("window_attention_bwd_tc: workspace too small or not 1024-byte aligned");
    return EMIP_ENOMEM;
  }
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  float* dsum = reinterpret_cast<float*>(static_cast<char*>(workspace) + ws_bytes_groups(B, g, ng));
  if ((rc = attn_dsum(dout, (long long)h * w * KC, out, (long long)h * w * KC, dsum, B, h * w, EMIP_LAYOUT_NC, st))) return rc;
  for (int i = 0; i < ng; ++i) {
    const int n = g[i].n, nprob = g[i].nblk * B;
    char* base = static_cast<char*>(workspace);
    const size_t tok_bytes = match_tc_split_bytes(nprob, n, KC);
    SplitParams sp;
    sp.src[0] = q; sp.src[1] = k; sp.src[2] = v; sp.src[3] = dout;
    for (int j = 0; j < 4; ++j) sp.dst[j] = reinterpret_cast<__nv_bfloat16*>(base + j * tok_bytes);
    sp.B = B; sp.h = h; sp.w = w; sp.n = n; sp.kv_shift = 0;
    AttnWinMap wm = {};
    wm.enabled = 1; wm.B = B; wm.h = h; wm.w = w;
    for (int j = 0; j < MAXBLK; ++j) {
      sp.r0[j] = wm.r0[j] = j < g[i].nblk ? g[i].r0[j] : 0;
      sp.c0[j] = wm.c0[j] = j < g[i].nblk ? g[i].c0[j] : 0;
      sp.bw[j] = wm.bw[j] = j < g[i].nblk ? g[i].bw[j] : 1;
    }
    win_split_kernel<<<dim3((n + TOK - 1) / TOK, nprob, 4), 256, 0, st>>>(sp);
    EMIP_CHECK_LAUNCH("window_attention_bwd (split)");
    AttnBwdTcArgs a = {};
    a.lse = lse; a.dsum = dsum; a.nb = nprob; a.nr = n; a.nc = n; a.out_layout = EMIP_LAYOUT_NC; a.out_stride_b = 0;
    a.sqrt_c = sqrtf((float)KC); a.ksplit = 1; a.win = wm;
    a.mode = ATTN_BWD_ROW; a.x_split = sp.dst[0]; a.y_split = sp.dst[1]; a.g_split = sp.dst[3]; a.z_split = sp.dst[2]; a.out = dq;
    if ((rc = attn_bwd_tc(a, st))) return rc;
    a.mode = ATTN_BWD_COL; a.x_split = sp.dst[1]; a.y_split = sp.dst[0]; a.g_split = sp.dst[2]; a.z_split = sp.dst[3]; a.out = dk;
    if ((rc = attn_bwd_tc(a, st))) return rc;
    a.mode = ATTN_BWD_PV; a.g_split = nullptr; a.out = dv;
    if ((rc = attn_bwd_tc(a, st))) return rc;
  }
  return EMIP_OK;
}


// =====================================================================================================================
// The whole FeatureTransformer (reference .../gmflow/transformer.py:433-482: num_layers blocks of a self-attention layer and a
// cross-attention + FFN layer, :349-401, :151-180) as ONE C-ABI call on token rows, inference.  Every tensor between two
// kernels lives in the form its consumer reads -- bf16 hi | lo operands written by the producer's epilogue:
//   * q / k / v projections: one GEMM per source tensor (self: q|k|v, N = 384; cross: q and k|v) whose epilogue writes the
//     rows straight into the WINDOW-ORDERED attention operands (rowmap: pixel -> slot; the cross layer's keys / values of
//     the other frame = an image shift in the same map) -- no fp32 q / k / v, no window gather, no split pass;
//   * the attention epilogue writes its output rows (scattered back to pixel order) as the pre-split A operand of `merge`;
//   * merge + norm1 (+ source) and mlp[2] + norm2 + source run LayerNorm in the GEMM epilogue and emit fp32 (residual for
//     later) AND the hi | lo rows of the next GEMM; the cross layer's normalised message goes directly into columns
//     128..255 of the MLP's [L][256] operand next to the source rows -- torch.cat([source, message]) is never made;
//   * mlp[0]'s epilogue applies the exact GELU and writes the hidden rows as the operand of mlp[2].
// Weights are split once per parameter version (emip_feature_transformer_prepare).
namespace {
constexpr int FT_W_PER_BLOCK = 16;     // self: q k v merge n1.w n1.b | cross: q k v merge n1.w n1.b mlp0 mlp2 n2.w n2.b
constexpr int FT_HID = 1024;

struct FtPrepOff { size_t self_qkv, self_merge, cross_q, cross_kv, cross_merge, mlp0, mlp2, total; };
FtPrepOff ft_prep_off() {
  FtPrepOff o;
  size_t p = 0;
  o.self_qkv = p; p += gemm_tc_split_b_bytes(3 * KC, KC);
  o.self_merge = p; p += gemm_tc_split_b_bytes(KC, KC);
  o.cross_q = p; p += gemm_tc_split_b_bytes(KC, KC);
  o.cross_kv = p; p += gemm_tc_split_b_bytes(2 * KC, KC);
  o.cross_merge = p; p += gemm_tc_split_b_bytes(KC, KC);
  o.mlp0 = p; p += gemm_tc_split_b_bytes(FT_HID, 2 * KC);
  o.mlp2 = p; p += gemm_tc_split_b_bytes(KC, FT_HID);
  o.total = p;
  return o;
}

struct FtWs {
  __nv_bfloat16 *xs_hi[2], *xs_lo[2];   // [L][256]: source rows | normalised message
  float* xf[2];                         // [L][128] fp32
  __nv_bfloat16 *q, *k, *v;             // window-ordered [rows][256 = hi | lo]
  __nv_bfloat16 *msg_hi, *msg_lo;       // [L][128]
  __nv_bfloat16 *hid_hi, *hid_lo;       // [L][1024]
  int2* rowmap[2];                      // unshifted / shifted
  size_t total;
};
FtWs ft_carve(void* base, size_t L, int npix) {
  FtWs w;
  char* p = static_cast<char*>(base);
  auto take = [&](size_t bytes) { char* r = p; p += emip_align_up(bytes, 1024); return r; };
  for (int i = 0; i < 2; ++i) {
    w.xs_hi[i] = reinterpret_cast<__nv_bfloat16*>(take(L * 256 * 2));
    w.xs_lo[i] = reinterpret_cast<__nv_bfloat16*>(take(L * 256 * 2));
    w.xf[i] = reinterpret_cast<float*>(take(L * KC * 4));
  }
  w.q = reinterpret_cast<__nv_bfloat16*>(take(L * 256 * 2));
  w.k = reinterpret_cast<__nv_bfloat16*>(take(L * 256 * 2));
  w.v = reinterpret_cast<__nv_bfloat16*>(take(L * 256 * 2));
  w.msg_hi = reinterpret_cast<__nv_bfloat16*>(take(L * KC * 2));
  w.msg_lo = reinterpret_cast<__nv_bfloat16*>(take(L * KC * 2));
  w.hid_hi = reinterpret_cast<__nv_bfloat16*>(take(L * FT_HID * 2));
  w.hid_lo = reinterpret_cast<__nv_bfloat16*>(take(L * FT_HID * 2));
  for (int i = 0; i < 2; ++i) w.rowmap[i] = reinterpret_cast<int2*>(take((size_t)npix * sizeof(int2)));
  w.total = (size_t)(p - static_cast<char*>(base));
  return w;
}

struct RowmapParams {
  int ng, B, w, npix;
  int n[4], nblk[4], base[4];
  int r0[4][MAXBLK], c0[4][MAXBLK], bw[4][MAXBLK];
};
// rowmap[pixel] = (slot of image 0's row, rows per image step): the block that contains the pixel, its group's base
__global__ void __launch_bounds__(256)
ft_rowmap_kernel(const RowmapParams rp, int2* __restrict__ rowmap) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= rp.npix) return;
  const int y = pix / rp.w, x = pix - y * rp.w;
  for (int g = 0; g < rp.ng; ++g)
    for (int b = 0; b < rp.nblk[g]; ++b) {
      const int bw = rp.bw[g][b], bh = rp.n[g] / bw, r0 = rp.r0[g][b], c0 = rp.c0[g][b];
      if (y >= r0 && y < r0 + bh && x >= c0 && x < c0 + bw) {
        rowmap[pix] = make_int2(rp.base[g] + b * rp.B * rp.n[g] + (y - r0) * bw + (x - c0), rp.n[g]);
        return;
      }
    }
}

// x [L][128] fp32 -> hi / lo rows of pitch ld (the first block's source operand)
__global__ void __launch_bounds__(256)
ft_split_rows_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long long L, int ld) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // one float4 per thread
  if (i >= L * 32) return;
  const long long r = i >> 5;
  const int c4 = (int)(i & 31);
  const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * KC) + c4);
  const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
  const uint32_t u0 = *reinterpret_cast<const uint32_t*>(&h0), u1 = *reinterpret_cast<const uint32_t*>(&h1);
  const __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - __uint_as_float(u0 << 16), v.y - __uint_as_float(u0 & 0xffff0000u));
  const __nv_bfloat162 l1 = __floats2bfloat162_rn(v.z - __uint_as_float(u1 << 16), v.w - __uint_as_float(u1 & 0xffff0000u));
  *reinterpret_cast<uint2*>(hi + r * ld + 4 * c4) = make_uint2(u0, u1);
  *reinterpret_cast<uint2*>(lo + r * ld + 4 * c4) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
}

struct FtGeom {
  Group g[4];
  int ng;
  int base[4];
};

int ft_geom(int B, int h, int w, int splits, int shift, FtGeom* out) {
  Group tmp[16];
  const int ng = make_groups(h, w, splits, shift, tmp, 16);
  if (ng < 0 || ng > 4) return -1;
  out->ng = ng;
  int base = 0;
  for (int i = 0; i < ng; ++i) {
    out->g[i] = tmp[i];
    out->base[i] = base;
    base += tmp[i].nblk * B * tmp[i].n;
  }
  return base == B * h * w ? 0 : -1;
}

// one attention layer on the window-ordered operands: per block group one fused flash launch, rows scattered to pixel order
int ft_attention(const FtWs& ws, const FtGeom& ge, int B, int h, int w, cudaStream_t st) {
  // ONE launch for all block groups of the layer (attn_tc.cuh: AttnTcArgs::more).  The groups of a shifted layer (22 x 22, 22 x 11 /
  // 11 x 22 and 11 x 11 token blocks) were three launches, each ending in its own partial wave: at 8 pairs 64 + 128 + 64 items
  // for 148 persistent CTAs.  Sets ordered by token count, largest first: items are dealt round-robin, heavy ones first.
  if (ge.ng < 1 || ge.ng > 3) { emip_set_error("feature_transformer: %d block groups", ge.ng); return EMIP_ENOSYS; }
  int order[3] = {0, 1, 2};
  for (int i = 0; i < ge.ng; ++i)
    for (int j = i + 1; j < ge.ng; ++j)
      if (ge.g[order[j]].n > ge.g[order[i]].n) { const int t = order[i]; order[i] = order[j]; order[j] = t; }
  AttnTcArgs a = {};
  a.win.enabled = 1; a.win.B = B; a.win.h = h; a.win.w = w;
  int blk = 0;
  for (int s = 0; s < ge.ng; ++s) {
    const Group& g = ge.g[order[s]];
    if (blk + g.nblk > 16) { emip_set_error("feature_transformer: more than 16 attention blocks"); return EMIP_ENOSYS; }
    for (int j = 0; j < g.nblk; ++j) { a.win.r0[blk + j] = g.r0[j]; a.win.c0[blk + j] = g.c0[j]; a.win.bw[blk + j] = g.bw[j]; }
    const size_t off = (size_t)ge.base[order[s]] * 256;
    if (s == 0) {
      a.q_split = ws.q + off; a.k_split = ws.k + off; a.v_split = ws.v + off;
      a.nb = g.nblk * B; a.nq = g.n; a.nk = g.n;
    } else {
      AttnTcArgs::More& m = a.more[s - 1];
      m.q_split = ws.q + off; m.k_split = ws.k + off; m.v_split = ws.v + off;
      m.nb = g.nblk * B; m.nq = g.n; m.nk = g.n; m.blk0 = blk;
    }
    blk += g.nblk;
  }
  for (int j = blk; j < 16; ++j) a.win.bw[j] = 1;
  a.n_more = ge.ng - 1;
  a.out = nullptr; a.out_stride_b = 0; a.lse = nullptr;
  a.out_hi = ws.msg_hi; a.out_lo = ws.msg_lo; a.out_ld = KC;
  a.out_layout = EMIP_LAYOUT_NC;
  a.sqrt_c = sqrtf((float)KC);
  a.ksplit = 1;
  return attn_tc_fwd(a, st);
}
}  // namespace

extern "C" size_t emip_feature_transformer_weight_bytes(int n_blocks) {
  return n_blocks > 0 ? (size_t)n_blocks * ft_prep_off().total : 0;
}

// weights: n_blocks x 16 pointers in the reference's state_dict order of one TransformerBlock (transformer.py:349-375):
//   self_attn.{q_proj, k_proj, v_proj, merge}.weight [128,128], self_attn.norm1.{weight, bias},
//   cross_attn_ffn.{q_proj, k_proj, v_proj, merge}.weight, .norm1.{weight, bias}, .mlp.0.weight [1024,256], .mlp.2.weight [128,1024],
//   .norm2.{weight, bias}.  The LayerNorm vectors are read in place by the forward; only the matrices are prepared.
extern "C" int emip_feature_transformer_prepare(const float* const* weights, int n_blocks, void* prep, void* stream) {
  EMIP_CHECK_ARG(weights && prep && n_blocks > 0, "feature_transformer_prepare: bad arguments");
  EMIP_CHECK_ARG(reinterpret_cast<uintptr_t>(prep) % 1024 == 0, "feature_transformer_prepare: prep must be 1024-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const FtPrepOff o = ft_prep_off();
  for (int b = 0; b < n_blocks; ++b) {
    const float* const* w = weights + (size_t)b * FT_W_PER_BLOCK;
    for (int i = 0; i < FT_W_PER_BLOCK; ++i) EMIP_CHECK_ARG(w[i] != nullptr, "feature_transformer_prepare: null weight %d of block %d", i, b);
    char* pb = static_cast<char*>(prep) + (size_t)b * o.total;
    int rc;
    const size_t row = gemm_tc_split_b_bytes(KC, KC);             // one [128][2 * 128] matrix
    for (int i = 0; i < 3; ++i)
      if ((rc = gemm_tc_split_b(w[i], KC, KC, KC, pb + o.self_qkv + i * row, st))) return rc;
    if ((rc = gemm_tc_split_b(w[3], KC, KC, KC, pb + o.self_merge, st))) return rc;
    if ((rc = gemm_tc_split_b(w[6], KC, KC, KC, pb + o.cross_q, st))) return rc;
    for (int i = 0; i < 2; ++i)
      if ((rc = gemm_tc_split_b(w[7 + i], KC, KC, KC, pb + o.cross_kv + i * row, st))) return rc;
    if ((rc = gemm_tc_split_b(w[9], KC, KC, KC, pb + o.cross_merge, st))) return rc;
    if ((rc = gemm_tc_split_b(w[12], 2 * KC, FT_HID, 2 * KC, pb + o.mlp0, st))) return rc;
    if ((rc = gemm_tc_split_b(w[13], FT_HID, KC, FT_HID, pb + o.mlp2, st))) return rc;
  }
  return EMIP_OK;
}

extern "C" size_t emip_feature_transformer_workspace(int B, int h, int w, int C) {
  if (B <= 0 || h <= 0 || w <= 0 || C != KC) return 0;
  return ft_carve(nullptr, (size_t)B * h * w, h * w).total;
}

// x, out [B][h*w][128] token rows, B = 2 x pairs (frame 1 | frame 2 stacked on the batch axis, transformer.py:461); out may
// alias x.  Block i uses shifted windows when i is odd (transformer.py:425).
extern "C" int emip_feature_transformer_fwd(const float* x, float* out, const float* const* weights, const void* prep, int n_blocks,
                                            void* workspace, size_t ws_bytes, int B, int h, int w, int C, int num_splits,
                                            float eps, void* stream) {
  return emip_feature_transformer_fwd_ex(x, out, nullptr, weights, prep, n_blocks, workspace, ws_bytes, B, h, w, C, num_splits, eps, stream);
}

// out_split (optional): the output rows once more as bf16 [B][h*w][hi 128 | lo 128] -- the operand format of the matching / flow
// attention kernels and of the token-row GEMMs -- written by the last block's LayerNorm epilogue (no split pass over `out`)
extern "C" int emip_feature_transformer_fwd_ex(const float* x, float* out, void* out_split, const float* const* weights, const void* prep,
                                               int n_blocks, void* workspace, size_t ws_bytes, int B, int h, int w, int C, int num_splits,
                                               float eps, void* stream) {
  if (B == 0 || n_blocks == 0) return EMIP_OK;
  EMIP_CHECK_ARG(reinterpret_cast<uintptr_t>(out_split) % 128 == 0, "feature_transformer_fwd: out_split must be 128-byte aligned");
  EMIP_CHECK_ARG(x && out && weights && prep && workspace, "feature_transformer_fwd: null pointer");
  EMIP_CHECK_ARG(B > 0 && B % 2 == 0 && h > 0 && w > 0 && n_blocks > 0, "feature_transformer_fwd: bad shape B=%d (even) h=%d w=%d", B, h, w);
  if (C != KC) { emip_set_error("feature_transformer_fwd: C=%d unsupported (kernels are built for the model's C=128)", C); return EMIP_ENOSYS; }
  const int npix = h * w;
  const long long L = (long long)B * npix;
  EMIP_CHECK_ARG(L < 0x7fffffffLL / 2, "feature_transformer_fwd: too many rows");
  FtGeom geom[2];
  if (num_splits < 2 || ft_geom(B, h, w, num_splits, 0, &geom[0]) || ft_geom(B, h, w, num_splits, 1, &geom[1])) {
    emip_set_error("feature_transformer_fwd: unsupported window geometry h=%d w=%d num_splits=%d", h, w, num_splits);
    return EMIP_ENOSYS;
  }
  if (ws_bytes < emip_feature_transformer_workspace(B, h, w, C) || reinterpret_cast<uintptr_t>(workspace) % 1024 != 0) {
    emip_set_error("feature_transformer_fwd: workspace too small or not 1024-byte aligned");
    return EMIP_ENOMEM;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const FtWs ws = ft_carve(workspace, (size_t)L, npix);
  const FtPrepOff po = ft_prep_off();
  int rc;
  for (int s = 0; s < 2; ++s) {
    RowmapParams rp = {};
    rp.ng = geom[s].ng; rp.B = B; rp.w = w; rp.npix = npix;
    for (int g = 0; g < geom[s].ng; ++g) {
      rp.n[g] = geom[s].g[g].n; rp.nblk[g] = geom[s].g[g].nblk; rp.base[g] = geom[s].base[g];
      for (int j = 0; j < MAXBLK; ++j) { rp.r0[g][j] = geom[s].g[g].r0[j]; rp.c0[g][j] = geom[s].g[g].c0[j]; rp.bw[g][j] = geom[s].g[g].bw[j] > 0 ? geom[s].g[g].bw[j] : 1; }
    }
    ft_rowmap_kernel<<<(npix + 255) / 256, 256, 0, st>>>(rp, ws.rowmap[s]);
    EMIP_CHECK_LAUNCH("feature_transformer_fwd (rowmap)");
  }
  ft_split_rows_kernel<<<(unsigned)((L * 32 + 255) / 256), 256, 0, st>>>(x, ws.xs_hi[0], ws.xs_lo[0], L, 256);
  EMIP_CHECK_LAUNCH("feature_transformer_fwd (split)");

  auto base_gemm = [&](const __nv_bfloat16* a_hi, const __nv_bfloat16* a_lo, int a_ld, int K, const void* b_pre, int Nout) {
    GemmNT t = {};
    t.B = 1; t.M = (int)L; t.K = Nout; t.N = K;
    t.a_hi_pre = a_hi; t.a_lo_pre = a_lo; t.a_ld_pre = a_ld;
    t.b_pre = b_pre;
    t.ldc = Nout;
    return t;
  };
  for (int blk = 0; blk < n_blocks; ++blk) {
    const float* const* wt = weights + (size_t)blk * FT_W_PER_BLOCK;
    const char* pb = static_cast<const char*>(prep) + (size_t)blk * po.total;
    const int sh = blk & 1;
    const float* xin_f = blk == 0 ? x : ws.xf[0];
    float* xnext_f = blk == n_blocks - 1 ? out : ws.xf[0];
    // ---- self-attention layer (transformer.py:388-393)
    {
      GemmNT t = base_gemm(ws.xs_hi[0], ws.xs_lo[0], 256, KC, pb + po.self_qkv, 3 * KC);       // q | k | v of the source rows
      t.split_dst[0] = ws.q; t.split_dst[1] = ws.k; t.split_dst[2] = ws.v;
      t.rowmap = ws.rowmap[sh]; t.npix = npix; t.n_img = B; t.img_shift = 0;
      if ((rc = gemm_nt_tc(t, nullptr, 0, st, 1))) return rc;
    }
    if ((rc = ft_attention(ws, geom[sh], B, h, w, st))) return rc;
    {
      GemmNT t = base_gemm(ws.msg_hi, ws.msg_lo, KC, KC, pb + po.self_merge, KC);               // merge + norm1 + source (:171-172, :180)
      t.ln_gamma = wt[4]; t.ln_beta = wt[5]; t.ln_eps = eps; t.c_res = xin_f;
      t.c = ws.xf[1];
      t.ln_hi = ws.xs_hi[1]; t.ln_lo = ws.xs_lo[1]; t.ln_ld = 256;
      if ((rc = gemm_nt_tc(t, nullptr, 0, st, 1))) return rc;
    }
    // ---- cross-attention + FFN layer (transformer.py:396-401): keys / values from the block's INPUT rows of the other frame
    {
      GemmNT t = base_gemm(ws.xs_hi[1], ws.xs_lo[1], 256, KC, pb + po.cross_q, KC);
      t.split_dst[0] = ws.q;
      t.rowmap = ws.rowmap[sh]; t.npix = npix; t.n_img = B; t.img_shift = 0;
      if ((rc = gemm_nt_tc(t, nullptr, 0, st, 1))) return rc;
    }
    {
      GemmNT t = base_gemm(ws.xs_hi[0], ws.xs_lo[0], 256, KC, pb + po.cross_kv, 2 * KC);
      t.split_dst[0] = ws.k; t.split_dst[1] = ws.v;
      t.rowmap = ws.rowmap[sh]; t.npix = npix; t.n_img = B; t.img_shift = B / 2;      // image i's k, v feed image (i + B/2) % B
      if ((rc = gemm_nt_tc(t, nullptr, 0, st, 1))) return rc;
    }
    if ((rc = ft_attention(ws, geom[sh], B, h, w, st))) return rc;
    {
      GemmNT t = base_gemm(ws.msg_hi, ws.msg_lo, KC, KC, pb + po.cross_merge, KC);              // merge + norm1 -> message columns (:171-172)
      t.ln_gamma = wt[10]; t.ln_beta = wt[11]; t.ln_eps = eps; t.c_res = nullptr;
      t.c = nullptr;
      t.ln_hi = ws.xs_hi[1] + KC; t.ln_lo = ws.xs_lo[1] + KC; t.ln_ld = 256;
      if ((rc = gemm_nt_tc(t, nullptr, 0, st, 1))) return rc;
    }
    const bool last_split = blk == n_blocks - 1 && out_split != nullptr;
    MlpFusedArgs f = {};                                  // mlp + norm2 + source in ONE kernel, the hidden rows stay in TMEM (mlp_fused.cu)
    f.x_hi = ws.xs_hi[1]; f.x_lo = ws.xs_lo[1]; f.ldx = 256;
    f.w1 = pb + po.mlp0; f.w2 = pb + po.mlp2; f.L = (int)L; f.hid = FT_HID;
    f.gamma = wt[14]; f.beta = wt[15]; f.eps = eps; f.res = ws.xf[1]; f.ldr = KC;
    f.y = xnext_f; f.ldy = KC;
    f.out_hi = last_split ? out_split : static_cast<void*>(ws.xs_hi[0]);
    f.out_lo = last_split ? static_cast<void*>(static_cast<__nv_bfloat16*>(out_split) + KC) : static_cast<void*>(ws.xs_lo[0]);
    f.out_ld = 256;
    if (!gemm_tc_debug_two_launch_mlp() && mlp_fused_supported(f)) {
      if ((rc = mlp_fused_tc(f, st))) return rc;
      continue;
    }
    {
      GemmNT t = base_gemm(ws.xs_hi[1], ws.xs_lo[1], 256, 2 * KC, pb + po.mlp0, FT_HID);        // mlp[0] on [source | message] + GELU (:175)
      t.c_hi = ws.hid_hi; t.c_lo = ws.hid_lo; t.ldc_split = FT_HID; t.c_act = 1;
      if ((rc = gemm_nt_tc(t, nullptr, 0, st, 1))) return rc;
    }
    {
      GemmNT t = base_gemm(ws.hid_hi, ws.hid_lo, FT_HID, FT_HID, pb + po.mlp2, KC);             // mlp[2] + norm2 + source (:175-176, :180)
      t.ln_gamma = wt[14]; t.ln_beta = wt[15]; t.ln_eps = eps; t.c_res = ws.xf[1];
      t.c = xnext_f;
      t.ln_hi = ws.xs_hi[0]; t.ln_lo = ws.xs_lo[0]; t.ln_ld = 256;
      if (blk == n_blocks - 1 && out_split != nullptr) {
        t.ln_hi = out_split; t.ln_lo = static_cast<__nv_bfloat16*>(out_split) + KC;
      }
      if ((rc = gemm_nt_tc(t, nullptr, 0, st, 1))) return rc;
    }
  }
  return EMIP_OK;
}
