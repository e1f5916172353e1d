// a2: flow-propagation self-attention core -- C-ABI front end.
// Replaces reference model/EMIP_short/motion/gmflow/transformer.py:526-532
// (scores = q k^T / sqrt(C); softmax; prob @ flow) and its autograd backward
// (value = flow.detach(), gmflow.py:137, so only dq and dk exist).  The two
// projections (transformer.py:523-524) are linear_cn.cu; q and k then arrive channel-major (EMIP_FLAG_CHANNEL_MAJOR).
#include "common.cuh"
#include "../../include/emip_b200.h"
#include "pair_common.cuh"
#include "match_tc.cuh"
#include "pair_bwd_tc.cuh"
#include "gemm_tc.cuh"
#include <math.h>

namespace {
size_t d_bytes(int B, int N) { return emip_align_up(sizeof(float) * (size_t)B * N, 1024); }
}

extern "C" size_t emip_flow_attn_workspace(int B, int N, int C) {
  if (B < 0 || N <= 0 || C <= 0) return 0;
  return d_bytes(B, N) + 2 * match_tc_split_bytes(B, N, C) + pair_bwd_tc_chn_bytes(2 * B, N) + match_tc_streamk_bytes(B, N, N);
}

extern "C" int emip_flow_attn_fwd(const float* q, const float* k, const float* v, float* out, float* lse,
                                  void* workspace, size_t ws_bytes, int B, int N, int C, int flags, void* stream) {
  if (B == 0) return EMIP_OK;
  EMIP_CHECK_ARG(q && k && v && out, "flow_attn_fwd: null pointer");
  EMIP_CHECK_ARG(B >= 0 && N > 0, "flow_attn_fwd: bad shape B=%d N=%d", B, N);
  if (C != 128) {
    emip_set_error("flow_attn_fwd: C=%d unsupported (kernels are built for the model's C=128)", C);
    return EMIP_ENOSYS;
  }
  if (B == 0) return EMIP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int lay = (flags & EMIP_FLAG_CHANNEL_MAJOR) ? EMIP_LAYOUT_CN : EMIP_LAYOUT_NC;      // layout of q and k
  if (!(flags & EMIP_FLAG_EXACT_FP32) && match_tc_supported(N, N, C)) {
    size_t sb = match_tc_split_bytes(B, N, C);
    const size_t sk_off = d_bytes(B, N) + 2 * sb + pair_bwd_tc_chn_bytes(2 * B, N), sk_bytes = match_tc_streamk_bytes(B, N, N);
    if (workspace == nullptr || ws_bytes < sk_off + sk_bytes || reinterpret_cast<uintptr_t>(workspace) % 1024) {
      emip_set_error("flow_attn_fwd: workspace too small or not 1024-byte aligned");
      return EMIP_ENOMEM;
    }
    char* p = static_cast<char*>(workspace) + d_bytes(B, N);
    int rc;
    if ((rc = match_tc_split(q, nullptr, p, B, N, C, lay, 0, st))) return rc;
    if ((rc = match_tc_split(k, nullptr, p + sb, B, N, C, lay, 0, st))) return rc;
    MatchTcArgs a = {};
    a.x_split = p; a.y_split = p + sb; a.nbx = B; a.nby = B;
    a.v = v; a.v_stride_b = 2LL * N; a.grid_w = 0; a.sub_grid = 0;
    a.terms = (flags & EMIP_FLAG_BF16) ? 1 : 3;
    a.out = out; a.lse = lse; a.nb = B; a.nq = N; a.nk = N; a.y_shift = 0; a.y_mod = B;
    a.s_out = nullptr; a.s_first = 0; a.s_count = 0;
    a.sqrt_c = sqrtf((float)C);
    a.sk_ws = static_cast<char*>(workspace) + sk_off; a.sk_bytes = sk_bytes;
    a.schedule = (flags & EMIP_FLAG_SCHED_STREAMK) ? 1 : (flags & EMIP_FLAG_SCHED_ITEMS) ? 2 : 0;
    return match_tc_fwd(a, st);
  }
  PairFwdArgs a = {};
  a.x = q; a.y = k; a.v = v; a.v_stride_b = 2LL * N; a.sub = nullptr;
  a.out = out; a.lse = lse; a.s_out = nullptr;
  a.nb = B; a.nq = N; a.nk = N; a.y_shift = 0;
  a.x_layout = a.y_layout = lay;
  a.sqrt_c = sqrtf((float)C);                       // transformer.py:528
  return pair_fwd_simt(a, st);
}

// a2 in one call from the transformer's pre-split token rows: see include/emip_b200.h
extern "C" size_t emip_flow_attn_tokens_workspace(int B, int N, int C) {
  if (B < 0 || N <= 0 || C != 128) return 0;
  return 2 * match_tc_split_bytes(B, N, C) + 2 * emip_align_up(gemm_nt_tc_scratch_bytes_presplit(C, C) + 1024, 1024) + match_tc_streamk_bytes(B, N, N);
}

extern "C" int emip_flow_attn_tokens_fwd(const void* x_split, const float* wq, const float* bq, const float* wk, const float* bk,
                                         const float* v, float* out, void* workspace, size_t ws_bytes, int B, int N, int C, int flags,
                                         void* stream) {
  if (B == 0) return EMIP_OK;
  EMIP_CHECK_ARG(x_split && wq && bq && wk && bk && v && out && workspace, "flow_attn_tokens_fwd: null pointer");
  EMIP_CHECK_ARG(B > 0 && N > 0, "flow_attn_tokens_fwd: bad shape B=%d N=%d", B, N);
  if (C != 128 || !match_tc_supported(N, N, C) || (flags & EMIP_FLAG_EXACT_FP32)) {
    emip_set_error("flow_attn_tokens_fwd: needs the tensor-core path (C=128, 16 <= N <= 2048, no EMIP_FLAG_EXACT_FP32)");
    return EMIP_ENOSYS;
  }
  EMIP_CHECK_ARG((long long)B * N < 0x7fffffffLL, "flow_attn_tokens_fwd: too many rows");
  if (ws_bytes < emip_flow_attn_tokens_workspace(B, N, C) || reinterpret_cast<uintptr_t>(workspace) % 1024 != 0) {
    emip_set_error("flow_attn_tokens_fwd: workspace too small or not 1024-byte aligned");
    return EMIP_ENOMEM;
  }
  EMIP_CHECK_ARG(reinterpret_cast<uintptr_t>(x_split) % 128 == 0 && ((reinterpret_cast<uintptr_t>(bq) | reinterpret_cast<uintptr_t>(bk)) & 15) == 0,
                 "flow_attn_tokens_fwd: x_split must be 128-byte, bq / bk 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t sb = match_tc_split_bytes(B, N, C), wb = emip_align_up(gemm_nt_tc_scratch_bytes_presplit(C, C) + 1024, 1024);
  char* base = static_cast<char*>(workspace);
  __nv_bfloat16* qs = reinterpret_cast<__nv_bfloat16*>(base);             // query rows [B*N][hi 128 | lo 128]
  __nv_bfloat16* ks = reinterpret_cast<__nv_bfloat16*>(base + sb);        // key rows
  const __nv_bfloat16* src = static_cast<const __nv_bfloat16*>(x_split);
  const float* w[2] = {wq, wk};
  const float* bias[2] = {bq, bk};
  __nv_bfloat16* dst[2] = {qs, ks};
  for (int i = 0; i < 2; ++i) {                                            // transformer.py:523-524: key = k_proj(q_proj(x))
    GemmNT t = {};
    t.B = 1; t.M = B * N; t.K = C; t.N = C;
    t.a_hi_pre = src; t.a_lo_pre = src + C; t.a_ld_pre = 2 * C;
    t.bm = w[i]; t.ldb = C;
    t.c_hi = dst[i]; t.c_lo = dst[i] + C; t.ldc_split = 2 * C;
    t.c_bias = bias[i];
    t.ldc = C;
    if (int rc = gemm_nt_tc(t, base + 2 * sb + i * wb, wb, st, 1)) return rc;
    src = dst[i];
  }
  MatchTcArgs a = {};
  a.x_split = qs; a.y_split = ks; a.nbx = B; a.nby = B;
  a.v = v; a.v_stride_b = 2LL * N; a.grid_w = 0; a.sub_grid = 0;
  a.terms = (flags & EMIP_FLAG_BF16) ? 1 : 3;
  a.out = out; a.lse = nullptr; a.nb = B; a.nq = N; a.nk = N; a.y_shift = 0; a.y_mod = B;
  a.s_out = nullptr; a.s_first = 0; a.s_count = 0;
  a.sqrt_c = sqrtf((float)C);
  a.sk_ws = base + 2 * sb + 2 * wb; a.sk_bytes = match_tc_streamk_bytes(B, N, N);
  a.schedule = (flags & EMIP_FLAG_SCHED_STREAMK) ? 1 : (flags & EMIP_FLAG_SCHED_ITEMS) ? 2 : 0;
  return match_tc_fwd(a, st);
}

extern "C" int emip_flow_attn_bwd(const float* q, const float* k, const float* v, const float* out, const float* lse,
                                  const float* dout, float* dq, float* dk, void* workspace, size_t ws_bytes,
                                  int B, int N, int C, int flags, void* stream) {
  if (B == 0) return EMIP_OK;
  EMIP_CHECK_ARG(q && k && v && out && lse && dout && dq && dk, "flow_attn_bwd: null pointer");
  if (C != 128) {
    emip_set_error("flow_attn_bwd: C=%d unsupported (kernels are built for the model's C=128)", C);
    return EMIP_ENOSYS;
  }
  if (B == 0) return EMIP_OK;
  if (workspace == nullptr || ws_bytes < d_bytes(B, N)) {
    emip_set_error("flow_attn_bwd: workspace too small");
    return EMIP_ENOMEM;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int lay = (flags & EMIP_FLAG_CHANNEL_MAJOR) ? EMIP_LAYOUT_CN : EMIP_LAYOUT_NC;      // layout of q, k, dq, dk
  float* D = static_cast<float*>(workspace);      // D_i = dO_i . O_i
  int rc;
  if ((rc = launch_rowdot2(dout, out, nullptr, D, B, N, st))) return rc;
  if (!(flags & EMIP_FLAG_EXACT_FP32) && pair_bwd_tc_supported(N, N, C) && N % 8 == 0) {
    const size_t sb = match_tc_split_bytes(B, N, C);
    if (ws_bytes < d_bytes(B, N) + 2 * sb + pair_bwd_tc_chn_bytes(2 * B, N) || reinterpret_cast<uintptr_t>(workspace) % 1024) {
      emip_set_error("flow_attn_bwd: workspace too small or not 1024-byte aligned");
      return EMIP_ENOMEM;
    }
    char* tok = static_cast<char*>(workspace) + d_bytes(B, N);        // [q batches | k batches] token-major
    char* chn = tok + 2 * sb;
    if ((rc = match_tc_split(q, k, tok, B, N, C, lay, 0, st))) return rc;
    if ((rc = pair_bwd_tc_split_chn(q, k, chn, B, N, lay, st))) return rc;
    PairBwdTcArgs t = {};
    t.tok_split = tok; t.tok_split_y = tok; t.chn_split_y = chn; t.n_split = 2 * B;
    t.nb = B; t.nr = N; t.nc = N; t.dx_layout = lay; t.sqrt_c = sqrtf((float)C);
    // dq_i = sum_j P_ij (dO_i.v_j - D_i) k_j / sqrt(C)
    t.x_base = 0; t.y_base = B; t.dx = dq;
    t.l1 = lse; t.u = dout; t.u0 = D; t.t = v; t.t_stride_b = 2LL * N;
    if ((rc = pair_bwd_tc(t, st))) return rc;
    // dk_j = sum_i P_ij (dO_i.v_j - D_i) q_i / sqrt(C): rows j, softmax statistics live on the columns i
    t.x_base = B; t.y_base = 0; t.dx = dk;
    t.l1 = t.u = t.u0 = t.t = nullptr;
    t.l2 = lse; t.w = dout; t.w0 = D; t.t2 = v; t.t2_stride_b = 2LL * N;
    return pair_bwd_tc(t, st);
  }
  PairBwdArgs a = {};
  a.nb = B; a.nr = N; a.nc = N;
  a.x_layout = a.y_layout = a.dx_layout = lay;
  a.sqrt_c = sqrtf((float)C);
  // dq_i = sum_j P_ij (dO_i.v_j - D_i) k_j / sqrt(C)
  a.x = q; a.y = k; a.dx = dq;
  a.l1 = lse; a.u = dout; a.u0 = D; a.t = v; a.t_stride_b = 2LL * N;
  if ((rc = pair_bwd_simt(a, st))) return rc;
  // dk_j = sum_i P_ij (dO_i.v_j - D_i) q_i / sqrt(C): rows j, softmax statistics live on the columns i
  a.x = k; a.y = q; a.dx = dk;
  a.l1 = a.u = a.u0 = a.t = nullptr;
  a.l2 = lse; a.w = dout; a.w0 = D; a.t2 = v; a.t2_stride_b = 2LL * N;
  return pair_bwd_simt(a, st);
}
