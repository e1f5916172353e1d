// a5 forward on the tensor cores: EMIP_long space-time memory read, reference model/EMIP_long/LTM.py:49-68
//   p = softmax over the memory axis of (m_in^T q_in / sqrt(De));  mem = m_out p
// and f2, the attention core of the GMFlow FeatureTransformer -- both are  out = softmax(Q K^T / sqrt(128)) V  and run
// as ONE fused flash-style kernel (attn_tc.cu: S in TMEM, online softmax with a lazy reference, P back to TMEM as the
// A operand of a TS-mode UMMA; 3-term split-bf16 operands, fp32 accumulation) after three operand-split passes.
// The model calls the memory read with B = 1 (16 row tiles for 148 SMs): the key tiles are then dealt to `ksplit` CTAs
// per row tile and a small kernel merges the partial (O, m, l) states.
// The exact-fp32 CUDA-core path (memory_read.cu) stays as the reference path and does the backward.
#include "common.cuh"
#include "../../include/emip_b200.h"
#include "pair_common.cuh"
#include "match_tc.cuh"
#include "pair_bwd_tc.cuh"
#include "attn_tc.cuh"
#include <math.h>

namespace {
constexpr int KC = 128;
size_t alf(size_t n_floats) { return emip_align_up(n_floats * sizeof(float), 1024); }

int pick_ksplit(int B, int Q, int M) {
  const int nrt = (Q + 127) / 128, nkt = (M + 127) / 128;
  if (B <= 0 || nrt <= 0) return 1;
  int ks = emip_num_sms() / (B * nrt);                   // items x splits <= SMs: one wave, no second round for a few CTAs
  if (ks > nkt) ks = nkt;
  if (ks > 16) ks = 16;
  return ks < 1 ? 1 : ks;
}

struct Ws {
  char *qs, *ks, *vs;
  float* part;
  float2* ml;
};

// the values are split without a transpose either way: channel-major sources feed the K-major operand form,
// token-major sources the MN-major one
size_t carve(char* base, int B, int M, int Q, int layout, Ws* w) {
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += emip_align_up(bytes, 1024); return p; };
  const int ksp = pick_ksplit(B, Q, M);
  char* qs = take(match_tc_split_bytes(B, Q, KC));
  char* ks = take(match_tc_split_bytes(B, M, KC));
  char* vs = take(layout == EMIP_LAYOUT_CN ? pair_bwd_tc_chn_bytes(B, M) : match_tc_split_bytes(B, M, KC));
  float* part = reinterpret_cast<float*>(take(ksp > 1 ? alf((size_t)ksp * B * KC * Q) : 0));
  float2* ml = reinterpret_cast<float2*>(take(ksp > 1 ? alf((size_t)2 * ksp * B * Q) : 0));
  if (w) { w->qs = qs; w->ks = ks; w->vs = vs; w->part = part; w->ml = ml; }
  return off;
}
}  // namespace

// Shared core: out[b][q][:] (layout `out_layout`, batch stride out_stride_b) = softmax_m(Q K^T / sqrt(128)) V.
// q [B][..] has Q tokens, k / v have M tokens, all 128 channels, all in `layout` (CN = [128][n], NC = [n][128]).
static int attention_tc_core(const float* q_in, const float* m_in, const float* m_out, int layout, float* mem,
                             long long mem_stride_b, int out_layout, float* lse, void* workspace, int B, int M, int Q,
                             cudaStream_t st) {
  Ws w;
  carve(static_cast<char*>(workspace), B, M, Q, layout, &w);
  int rc;
  if ((rc = match_tc_split(q_in, nullptr, w.qs, B, Q, KC, layout, 0, st))) return rc;
  if ((rc = match_tc_split(m_in, nullptr, w.ks, B, M, KC, layout, 0, st))) return rc;
  if (layout == EMIP_LAYOUT_CN) rc = pair_bwd_tc_split_chn(m_out, nullptr, w.vs, B, M, layout, st);
  else rc = match_tc_split(m_out, nullptr, w.vs, B, M, KC, layout, 0, st);
  if (rc) return rc;
  AttnTcArgs a = {};
  a.q_split = w.qs; a.k_split = w.ks; a.v_split = w.vs; a.v_chn = layout == EMIP_LAYOUT_CN;
  a.out = mem; a.out_stride_b = mem_stride_b; a.lse = lse;
  a.part_o = w.part; a.part_ml = w.ml;
  a.nb = B; a.nq = Q; a.nk = M; a.out_layout = out_layout;
  a.sqrt_c = sqrtf((float)KC);
  a.ksplit = pick_ksplit(B, Q, M);
  if ((rc = attn_tc_fwd(a, st))) return rc;
  return attn_tc_merge(a, st);
}

extern "C" size_t emip_memory_read_tc_workspace(int B, int De, int Do, int M, int Q) {
  if (B < 0 || M <= 0 || Q <= 0 || De != KC || Do != KC) return 0;
  return carve(nullptr, B, M, Q, EMIP_LAYOUT_CN, nullptr);
}

extern "C" int emip_memory_read_fwd_tc(const float* m_in, const float* m_out, const float* q_in, float* mem,
                                       long long mem_stride_b, float* lse, void* workspace, size_t ws_bytes, int B, int De,
                                       int Do, int M, int Q, void* stream) {
  if (B == 0) return EMIP_OK;
  EMIP_CHECK_ARG(m_in && m_out && q_in && mem && workspace, "memory_read_fwd_tc: null pointer");
  EMIP_CHECK_ARG(M >= 16 && Q > 0, "memory_read_fwd_tc: bad shape M=%d Q=%d", M, Q);
  if (De != KC || Do != KC) {
    emip_set_error("memory_read_fwd_tc: De=%d Do=%d unsupported (kernels are built for the model's 128/128)", De, Do);
    return EMIP_ENOSYS;
  }
  if (ws_bytes < emip_memory_read_tc_workspace(B, De, Do, M, Q) || reinterpret_cast<uintptr_t>(workspace) % 1024 != 0) {
    emip_set_error("memory_read_fwd_tc: workspace too small or not 1024-byte aligned");
    return EMIP_ENOMEM;
  }
  return attention_tc_core(q_in, m_in, m_out, EMIP_LAYOUT_CN, mem, mem_stride_b, EMIP_LAYOUT_CN, lse, workspace, B, M, Q,
                           (cudaStream_t)stream);
}

// f2: the attention core of the GMFlow FeatureTransformer, reference
// model/EMIP_short/motion/gmflow/transformer.py:8-16 (single_head_full_attention) and :46-105 (the per-window attention of
// single_head_split_window_attention): out = softmax(q k^T / sqrt(C)) v for nb independent problems of n tokens,
// q, k, v, out token-major [nb][n][128].
extern "C" size_t emip_attention_tc_workspace(int nb, int n, int C) {
  if (nb < 0 || n <= 0 || C != KC) return 0;
  return carve(nullptr, nb, n, n, EMIP_LAYOUT_NC, nullptr);
}

extern "C" int emip_attention_fwd_tc(const float* q, const float* k, const float* v, float* out, void* workspace,
                                     size_t ws_bytes, int nb, int n, int C, void* stream) {
  if (nb == 0) return EMIP_OK;
  EMIP_CHECK_ARG(q && k && v && out && workspace, "attention_fwd_tc: null pointer");
  EMIP_CHECK_ARG(nb > 0 && n >= 1, "attention_fwd_tc: bad shape nb=%d n=%d", nb, n);
  if (C != KC) {
    emip_set_error("attention_fwd_tc: C=%d unsupported (kernels are built for the model's C=128)", C);
    return EMIP_ENOSYS;
  }
  if (ws_bytes < emip_attention_tc_workspace(nb, n, C) || reinterpret_cast<uintptr_t>(workspace) % 1024 != 0) {
    emip_set_error("attention_fwd_tc: workspace too small or not 1024-byte aligned");
    return EMIP_ENOMEM;
  }
  return attention_tc_core(q, k, v, EMIP_LAYOUT_NC, out, (long long)n * KC, EMIP_LAYOUT_NC, nullptr, workspace, nb, n, n,
                           (cudaStream_t)stream);
}

// ---- backward on the tensor cores (attn_bwd_tc.cu): three launches of one kernel (dQ; dK; dV) -----------------------
#include "attn_bwd_tc.cuh"

namespace {
struct BwdWs {
  char *qs, *ks, *vs, *dos;
  float *o, *lse, *dsum, *part;
};
// own_fwd: the forward is re-run here for O and the row log-sum-exp (f2 saves neither)
size_t carve_bwd(char* base, int B, int M, int Q, bool own_fwd, BwdWs* w) {
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += emip_align_up(bytes, 1024); return p; };
  BwdWs t;
  t.qs = take(match_tc_split_bytes(B, Q, KC));
  t.ks = take(match_tc_split_bytes(B, M, KC));
  t.vs = take(match_tc_split_bytes(B, M, KC));
  t.dos = take(match_tc_split_bytes(B, Q, KC));
  t.o = reinterpret_cast<float*>(take(own_fwd ? alf((size_t)B * Q * KC) : 0));
  t.lse = reinterpret_cast<float*>(take(own_fwd ? alf((size_t)B * Q) : 0));
  t.dsum = reinterpret_cast<float*>(take(alf((size_t)B * Q)));
  const size_t p1 = (size_t)pick_ksplit(B, Q, M) * B * Q * KC, p2 = (size_t)pick_ksplit(B, M, Q) * B * M * KC;
  t.part = reinterpret_cast<float*>(take(alf(p1 > p2 ? p1 : p2)));
  if (w) *w = t;
  return off;
}

// Gradients of out = softmax(q k^T / sqrt(128)) v given the operand splits, lse and D; dq [B][Q x 128], dk, dv [B][M x 128]
// in `layout` (contiguous batches).
int attention_bwd_core(const BwdWs& w, const float* lse, float* dq, float* dk, float* dv, int layout, int B, int M, int Q,
                       cudaStream_t st) {
  int rc;
  AttnBwdTcArgs a = {};
  a.lse = lse; a.dsum = w.dsum; a.part = w.part; a.nb = B; a.out_layout = layout; a.sqrt_c = sqrtf((float)KC);
  // dQ: rows = queries
  a.mode = ATTN_BWD_ROW; a.x_split = w.qs; a.y_split = w.ks; a.g_split = w.dos; a.z_split = w.vs;
  a.nr = Q; a.nc = M; a.out = dq; a.out_stride_b = (long long)Q * KC; a.ksplit = pick_ksplit(B, Q, M);
  if ((rc = attn_bwd_tc(a, st)) || (rc = attn_bwd_tc_sum(a, st))) return rc;
  // dK, dV: rows = keys, the score tile is recomputed transposed
  a.mode = ATTN_BWD_COL; a.x_split = w.ks; a.y_split = w.qs; a.g_split = w.vs; a.z_split = w.dos;
  a.nr = M; a.nc = Q; a.out = dk; a.out_stride_b = (long long)M * KC; a.ksplit = pick_ksplit(B, M, Q);
  if ((rc = attn_bwd_tc(a, st)) || (rc = attn_bwd_tc_sum(a, st))) return rc;
  a.mode = ATTN_BWD_PV; a.g_split = nullptr; a.out = dv;
  if ((rc = attn_bwd_tc(a, st)) || (rc = attn_bwd_tc_sum(a, st))) return rc;
  return EMIP_OK;
}
}  // namespace

extern "C" size_t emip_memory_read_bwd_tc_workspace(int B, int De, int Do, int M, int Q) {
  if (B < 0 || M <= 0 || Q <= 0 || De != KC || Do != KC) return 0;
  return carve_bwd(nullptr, B, M, Q, false, nullptr);
}

extern "C" int emip_memory_read_bwd_tc(const float* m_in, const float* m_out, const float* q_in, const float* mem,
                                       long long mem_stride_b, const float* lse, const float* dmem, long long dmem_stride_b,
                                       float* dm_in, float* dm_out, float* dq_in, void* workspace, size_t ws_bytes, int B,
                                       int De, int Do, int M, int Q, void* stream) {
  if (B == 0) return EMIP_OK;
  EMIP_CHECK_ARG(m_in && m_out && q_in && mem && lse && dmem && dm_in && dm_out && dq_in && workspace,
                 "memory_read_bwd_tc: null pointer");
  EMIP_CHECK_ARG(M > 0 && Q > 0, "memory_read_bwd_tc: bad shape M=%d Q=%d", M, Q);
  if (De != KC || Do != KC) {
    emip_set_error("memory_read_bwd_tc: De=%d Do=%d unsupported (kernels are built for the model's 128/128)", De, Do);
    return EMIP_ENOSYS;
  }
  if (ws_bytes < emip_memory_read_bwd_tc_workspace(B, De, Do, M, Q) || reinterpret_cast<uintptr_t>(workspace) % 1024 != 0) {
    emip_set_error("memory_read_bwd_tc: workspace too small or not 1024-byte aligned");
    return EMIP_ENOMEM;
  }
  cudaStream_t st = (cudaStream_t)stream;
  BwdWs w;
  carve_bwd(static_cast<char*>(workspace), B, M, Q, false, &w);
  int rc;
  if ((rc = match_tc_split(q_in, nullptr, w.qs, B, Q, KC, EMIP_LAYOUT_CN, 0, st))) return rc;
  if ((rc = match_tc_split(m_in, nullptr, w.ks, B, M, KC, EMIP_LAYOUT_CN, 0, st))) return rc;
  if ((rc = match_tc_split(m_out, nullptr, w.vs, B, M, KC, EMIP_LAYOUT_CN, 0, st))) return rc;
  for (int b = 0; b < B; ++b)      // dmem is the first half of a [B,256,H,W] gradient: explicit batch stride
    if ((rc = match_tc_split(dmem + (size_t)b * dmem_stride_b, nullptr, w.dos, 1, Q, KC, EMIP_LAYOUT_CN, b, st))) return rc;
  if ((rc = attn_dsum(dmem, dmem_stride_b, mem, mem_stride_b, w.dsum, B, Q, EMIP_LAYOUT_CN, st))) return rc;
  return attention_bwd_core(w, lse, dq_in, dm_in, dm_out, EMIP_LAYOUT_CN, B, M, Q, st);
}

extern "C" size_t emip_attention_bwd_tc_workspace(int nb, int n, int C) {
  if (nb < 0 || n <= 0 || C != KC) return 0;
  return carve_bwd(nullptr, nb, n, n, true, nullptr);
}

// f2 backward: dq, dk, dv of out = softmax(q k^T / sqrt(C)) v for nb problems of n tokens, all token-major [nb][n][C].
// Self-contained: re-runs the fused forward for O and the row log-sum-exp, then the three gradient launches.
extern "C" int emip_attention_bwd_tc(const float* q, const float* k, const float* v, const float* dout, float* dq, float* dk,
                                     float* dv, void* workspace, size_t ws_bytes, int nb, int n, int C, void* stream) {
  if (nb == 0) return EMIP_OK;
  EMIP_CHECK_ARG(q && k && v && dout && dq && dk && dv && workspace, "attention_bwd_tc: null pointer");
  EMIP_CHECK_ARG(nb > 0 && n >= 1, "attention_bwd_tc: bad shape nb=%d n=%d", nb, n);
  if (C != KC) {
    emip_set_error("attention_bwd_tc: C=%d unsupported (kernels are built for the model's C=128)", C);
    return EMIP_ENOSYS;
  }
  if (ws_bytes < emip_attention_bwd_tc_workspace(nb, n, C) || reinterpret_cast<uintptr_t>(workspace) % 1024 != 0) {
    emip_set_error("attention_bwd_tc: workspace too small or not 1024-byte aligned");
    return EMIP_ENOMEM;
  }
  cudaStream_t st = (cudaStream_t)stream;
  BwdWs w;
  carve_bwd(static_cast<char*>(workspace), nb, n, n, true, &w);
  int rc;
  if ((rc = match_tc_split(q, nullptr, w.qs, nb, n, KC, EMIP_LAYOUT_NC, 0, st))) return rc;
  if ((rc = match_tc_split(k, nullptr, w.ks, nb, n, KC, EMIP_LAYOUT_NC, 0, st))) return rc;
  if ((rc = match_tc_split(v, nullptr, w.vs, nb, n, KC, EMIP_LAYOUT_NC, 0, st))) return rc;
  if ((rc = match_tc_split(dout, nullptr, w.dos, nb, n, KC, EMIP_LAYOUT_NC, 0, st))) return rc;
  AttnTcArgs f = {};
  f.q_split = w.qs; f.k_split = w.ks; f.v_split = w.vs; f.v_chn = 0;
  f.out = w.o; f.out_stride_b = (long long)n * KC; f.lse = w.lse;
  f.nb = nb; f.nq = n; f.nk = n; f.out_layout = EMIP_LAYOUT_NC; f.sqrt_c = sqrtf((float)KC); f.ksplit = 1;
  if ((rc = attn_tc_fwd(f, st))) return rc;
  if ((rc = attn_dsum(dout, (long long)n * KC, w.o, (long long)n * KC, w.dsum, nb, n, EMIP_LAYOUT_NC, st))) return rc;
  return attention_bwd_core(w, w.lse, dq, dk, dv, EMIP_LAYOUT_NC, nb, n, n, st);
}
