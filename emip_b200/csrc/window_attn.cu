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
#include "attn_bwd_tc.cuh"
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
