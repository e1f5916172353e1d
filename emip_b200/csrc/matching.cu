// a1: GMFlow global matching -- C-ABI front end.
// Replaces reference model/EMIP_short/motion/gmflow/matching.py:8-41
// (global_correlation_softmax) and the autograd graph torch builds for it.
#include "common.cuh"
#include "../../include/emip_b200.h"
#include "pair_common.cuh"
#include "match_tc.cuh"
#include "pair_bwd_tc.cuh"
#include <math.h>

namespace {
struct MatchWs {
  float* grid;     // [2][N] pixel grid (geometry.py:5-21)
  float* u0;       // [nd*B][N]
  void* split;     // bf16 hi/lo token-major operands for the tensor-core paths
  size_t split_bytes;
  void* chn;       // bf16 hi/lo channel-major operands (tensor-core backward)
  void* sk;        // stream-K bookkeeping of the tensor-core forward
  size_t sk_bytes;
};

// uni- and bidirectional calls share one workspace: take the larger stream-K need
size_t sk_need(int B, int N) {
  const size_t a = match_tc_streamk_bytes(B, N, N), b = match_tc_streamk_bytes(2 * B, N, N);
  return a > b ? a : b;
}

size_t simt_bytes(int B, int N, int nd) { return emip_align_up(sizeof(float) * ((size_t)2 * N + (size_t)nd * B * N), 1024); }

int carve(void* ws, size_t ws_bytes, int B, int C, int N, int nd, MatchWs* out) {
  size_t need = simt_bytes(B, N, 2) + match_tc_split_bytes(2 * B, N, C) + pair_bwd_tc_chn_bytes(2 * B, N) +
                sk_need(B, N);
  if (ws == nullptr || ws_bytes < need) {
    emip_set_error("global_matching: workspace too small (%zu < %zu bytes)", ws_bytes, need);
    return EMIP_ENOMEM;
  }
  if (reinterpret_cast<uintptr_t>(ws) % 1024 != 0) {
    emip_set_error("global_matching: workspace must be 1024-byte aligned");
    return EMIP_EINVAL;
  }
  char* p = static_cast<char*>(ws);
  out->grid = reinterpret_cast<float*>(p);
  out->u0 = out->grid + 2 * (size_t)N;
  out->split = p + simt_bytes(B, N, 2);
  out->split_bytes = match_tc_split_bytes(2 * B, N, C);
  out->chn = p + simt_bytes(B, N, 2) + out->split_bytes;
  out->sk = static_cast<char*>(out->chn) + pair_bwd_tc_chn_bytes(2 * B, N);
  out->sk_bytes = sk_need(B, N);
  return EMIP_OK;
}
}  // namespace

extern "C" size_t emip_global_matching_workspace(int B, int C, int H, int W) {
  if (B < 0 || C <= 0 || H <= 0 || W <= 0) return 0;
  return simt_bytes(B, H * W, 2) + match_tc_split_bytes(2 * B, H * W, C) + pair_bwd_tc_chn_bytes(2 * B, H * W) +
         sk_need(B, H * W);
}

extern "C" int emip_global_matching_fwd(const float* f0, const float* f1, float* flow, float* corr, float* lse,
                                        void* workspace, size_t ws_bytes, int B, int C, int H, int W, int bidir,
                                        int flags, void* stream) {
  if (B == 0) return EMIP_OK;
  EMIP_CHECK_ARG(f0 && (f1 || (flags & EMIP_FLAG_PRESPLIT)) && flow, "global_matching_fwd: null pointer");
  EMIP_CHECK_ARG(B >= 0 && H > 0 && W > 0, "global_matching_fwd: bad shape B=%d H=%d W=%d", B, H, W);
  if (C != 128) {
    emip_set_error("global_matching_fwd: C=%d unsupported (kernels are built for the model's C=128)", C);
    return EMIP_ENOSYS;
  }
  if (B == 0) return EMIP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int N = H * W, nd = bidir ? 2 : 1;
  MatchWs ws;
  int rc = carve(workspace, ws_bytes, B, C, N, nd, &ws);
  if (rc) return rc;
  const bool reuse = (flags & EMIP_FLAG_REUSE_WORKSPACE) != 0;

  if (!(flags & EMIP_FLAG_EXACT_FP32) && match_tc_supported(N, N, C)) {
    // tensor-core path: one operand-split launch + one fused launch covering both directions
    const int layout = (flags & EMIP_FLAG_TOKEN_MAJOR) ? EMIP_LAYOUT_NC : EMIP_LAYOUT_CN;
    const bool presplit = (flags & EMIP_FLAG_PRESPLIT) != 0;      // f0 = the bf16 hi | lo operand of both frames
    if (presplit) EMIP_CHECK_ARG(reinterpret_cast<uintptr_t>(f0) % 16 == 0, "global_matching_fwd: the pre-split operand must be 16-byte aligned");
    if (!presplit && !reuse && (rc = match_tc_split(f0, f1, ws.split, B, N, C, layout, 0, st))) return rc;
    const void* split = presplit ? static_cast<const void*>(f0) : ws.split;
    MatchTcArgs a = {};
    a.x_split = split; a.y_split = split; a.nbx = 2 * B; a.nby = 2 * B;
    a.v = nullptr; a.v_stride_b = 0; a.grid_w = W; a.sub_grid = 1;   // analytic pixel grid (geometry.py:5-21)
    a.out = flow; a.lse = lse;
    a.nb = nd * B; a.nq = N; a.nk = N; a.y_shift = B; a.y_mod = 2 * B;
    a.s_out = corr; a.s_first = bidir ? B : 0; a.s_count = corr ? B : 0;
    a.sqrt_c = sqrtf((float)C);
    a.terms = (flags & EMIP_FLAG_BF16) ? 1 : 3;
    a.sk_ws = ws.sk; a.sk_bytes = ws.sk_bytes;
    a.schedule = (flags & EMIP_FLAG_SCHED_STREAMK) ? 1 : (flags & EMIP_FLAG_SCHED_ITEMS) ? 2 : 0;
    return match_tc_fwd(a, st);
  }
  if (flags & (EMIP_FLAG_TOKEN_MAJOR | EMIP_FLAG_PRESPLIT)) {
    emip_set_error("global_matching_fwd: EMIP_FLAG_TOKEN_MAJOR / EMIP_FLAG_PRESPLIT need the tensor-core path (C=128, 16 <= H*W <= 2048)");
    return EMIP_ENOSYS;
  }
  if ((rc = launch_coords_grid(ws.grid, H, W, st))) return rc;

  if (flags & EMIP_FLAG_BF16) {
    emip_set_error("global_matching_fwd: EMIP_FLAG_BF16 needs the tensor-core path (C=128, 16 <= H*W <= 2048)");
    return EMIP_ENOSYS;
  }
  PairFwdArgs a = {};
  a.v = ws.grid; a.v_stride_b = 0; a.sub = ws.grid;
  a.nb = B; a.nq = N; a.nk = N; a.y_shift = 0;
  a.x_layout = EMIP_LAYOUT_CN; a.y_layout = EMIP_LAYOUT_CN;
  a.sqrt_c = sqrtf((float)C);                       // matching.py:16  / (c ** 0.5)
  // forward direction: rows = f0 tokens, columns = f1 tokens
  a.x = f0; a.y = f1; a.out = flow; a.lse = lse;
  a.s_out = bidir ? nullptr : corr;                 // uni-directional: corr memory is S[b,i,j]
  if ((rc = pair_fwd_simt(a, st))) return rc;
  if (bidir) {
    // backward direction = the same problem on S^T (matching.py:29); its score tile is corr[b,j,i]
    a.x = f1; a.y = f0; a.out = flow + (size_t)B * 2 * N; a.lse = lse ? lse + (size_t)B * N : nullptr;
    a.s_out = corr;
    if ((rc = pair_fwd_simt(a, st))) return rc;
  }
  return EMIP_OK;
}

extern "C" int emip_global_matching_bwd(const float* f0, const float* f1, const float* flow, const float* lse,
                                        const float* dflow, const float* dcorr, float* df0, float* df1,
                                        void* workspace, size_t ws_bytes, int B, int C, int H, int W, int bidir,
                                        int flags, void* stream) {
  if (B == 0) return EMIP_OK;
  EMIP_CHECK_ARG(f0 && f1 && flow && lse && df0 && df1, "global_matching_bwd: null pointer");
  EMIP_CHECK_ARG(dflow || dcorr, "global_matching_bwd: both dflow and dcorr are NULL");
  if (C != 128) {
    emip_set_error("global_matching_bwd: C=%d unsupported (kernels are built for the model's C=128)", C);
    return EMIP_ENOSYS;
  }
  if (B == 0) return EMIP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int N = H * W, nd = bidir ? 2 : 1;
  MatchWs ws;
  int rc = carve(workspace, ws_bytes, B, C, N, nd, &ws);
  if (rc) return rc;
  if ((rc = launch_coords_grid(ws.grid, H, W, st))) return rc;
  if (dflow) {
    // u0_i = dflow_i . E_i  with  E_i = flow_i + g_i  (the softmax expectation)
    if ((rc = launch_rowdot2(dflow, flow, ws.grid, ws.u0, nd * B, N, st))) return rc;
  }
  const size_t hb = (size_t)B * N;   // offset of the backward-direction half
  if (!(flags & EMIP_FLAG_EXACT_FP32) && pair_bwd_tc_supported(N, N, C) && N % 8 == 0) {
    // tensor-core path: operands split once (token-major for S, channel-major for dX), one launch per gradient
    if ((rc = match_tc_split(f0, f1, ws.split, B, N, C, EMIP_LAYOUT_CN, 0, st))) return rc;
    if ((rc = pair_bwd_tc_split_chn(f0, f1, ws.chn, B, N, EMIP_LAYOUT_CN, st))) return rc;
    PairBwdTcArgs t = {};
    t.tok_split = ws.split; t.tok_split_y = ws.split; t.chn_split_y = ws.chn; t.n_split = 2 * B;
    t.nb = B; t.nr = N; t.nc = N; t.dx_layout = EMIP_LAYOUT_CN; t.sqrt_c = sqrtf((float)C);
    t.t = ws.grid; t.t_stride_b = 0; t.t2 = ws.grid; t.t2_stride_b = 0;
    t.e = dcorr; t.e_stride_b = (long long)N * N;
    // d f0 : rows i (f0), columns j (f1)
    t.x_base = 0; t.y_base = B; t.dx = df0;
    if (dflow) { t.l1 = lse; t.u = dflow; t.u0 = ws.u0; }
    if (dflow && bidir) { t.l2 = lse + hb; t.w = dflow + 2 * hb; t.w0 = ws.u0 + hb; }
    if (bidir) { t.e_stride_r = 1; t.e_stride_c = N; } else { t.e_stride_r = N; t.e_stride_c = 1; }
    if ((rc = pair_bwd_tc(t, st))) return rc;
    // d f1 : rows j (f1), columns i (f0) -- the same tile transposed
    t.x_base = B; t.y_base = 0; t.dx = df1;
    t.l1 = t.u = t.u0 = t.l2 = t.w = t.w0 = nullptr;
    if (dflow && bidir) { t.l1 = lse + hb; t.u = dflow + 2 * hb; t.u0 = ws.u0 + hb; }
    if (dflow) { t.l2 = lse; t.w = dflow; t.w0 = ws.u0; }
    if (bidir) { t.e_stride_r = N; t.e_stride_c = 1; } else { t.e_stride_r = 1; t.e_stride_c = N; }
    return pair_bwd_tc(t, st);
  }
  PairBwdArgs a = {};
  a.nb = B; a.nr = N; a.nc = N;
  a.x_layout = a.y_layout = a.dx_layout = EMIP_LAYOUT_CN;
  a.sqrt_c = sqrtf((float)C);
  a.t = ws.grid; a.t_stride_b = 0; a.t2 = ws.grid; a.t2_stride_b = 0;
  a.e = dcorr; a.e_stride_b = (long long)N * N;

  // d f0 : rows i (f0), columns j (f1)
  a.x = f0; a.y = f1; a.dx = df0;
  if (dflow) { a.l1 = lse; a.u = dflow; a.u0 = ws.u0; }
  if (dflow && bidir) { a.l2 = lse + hb; a.w = dflow + 2 * hb; a.w0 = ws.u0 + hb; }
  if (bidir) { a.e_stride_r = 1; a.e_stride_c = N; } else { a.e_stride_r = N; a.e_stride_c = 1; }
  if ((rc = pair_bwd_simt(a, st))) return rc;

  // d f1 : rows j (f1), columns i (f0) -- the same tile transposed
  a.x = f1; a.y = f0; a.dx = df1;
  a.l1 = a.u = a.u0 = a.l2 = a.w = a.w0 = nullptr;
  if (dflow && bidir) { a.l1 = lse + hb; a.u = dflow + 2 * hb; a.u0 = ws.u0 + hb; }
  if (dflow) { a.l2 = lse; a.w = dflow; a.w0 = ws.u0; }
  if (bidir) { a.e_stride_r = N; a.e_stride_c = 1; } else { a.e_stride_r = 1; a.e_stride_c = N; }
  return pair_bwd_simt(a, st);
}
