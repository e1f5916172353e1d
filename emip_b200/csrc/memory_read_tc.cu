// a5 forward on the tensor cores: EMIP_long space-time memory read, reference model/EMIP_long/LTM.py:49-68
//   p = softmax over the memory axis of (m_in^T q_in / sqrt(De));  mem = m_out p
// as two passes over the score matrix with kernels this library already has (both 3-term split-bf16, fp32 accumulate):
//   pass 1  match_tc_fwd (stream-K schedule, analytic-grid mode = no value table, any number of keys): the row
//           log-sum-exp L[q] of S = Q K^T / sqrt(De);
//   pass 2  pair_bwd_tc_kernel with its row term reduced to W = e^{S - L} (u = 0, u0 = -1) and the VALUES as the
//           channel-major column operand: out = W V -- S is recomputed, W goes to TMEM as bf16 hi|lo and feeds a
//           TS-mode UMMA.  Because W is already normalised by the final L, key splits simply add: the key tiles are
//           dealt to `ksplit` CTAs per row tile (the model calls this with B = 1: 16 row tiles for 148 SMs) and a
//           small kernel sums the partial outputs.
// The exact-fp32 CUDA-core path (memory_read.cu) stays as the reference path and does the backward.
#include "common.cuh"
#include "../../include/emip_b200.h"
#include "pair_common.cuh"
#include "match_tc.cuh"
#include "pair_bwd_tc.cuh"
#include <math.h>

namespace {
constexpr int KC = 128;
size_t alf(size_t n_floats) { return emip_align_up(n_floats * sizeof(float), 1024); }

int pick_ksplit(int B, int Q, int M) {
  const int nrt = (Q + 127) / 128, nkt = (M + 127) / 128;
  int ks = (emip_num_sms() + B * nrt - 1) / (B * nrt);
  if (ks > nkt) ks = nkt;
  if (ks > 16) ks = 16;
  return ks < 1 ? 1 : ks;
}

__global__ void fill_kernel(float* __restrict__ p, size_t n, float v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// mem[b][c][q] = scale * sum_s part[s][b][c][q]
__global__ void __launch_bounds__(256)
sum_parts_kernel(const float* __restrict__ part, float* __restrict__ mem, long long mem_stride_b, int ns, int B, int Q, float scale) {
  const int b = blockIdx.y;
  const size_t per = (size_t)KC * Q;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (size_t)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < ns; ++s) acc += __ldcs(part + ((size_t)s * B + b) * per + i);
    mem[(size_t)b * mem_stride_b + i] = acc * scale;
  }
}

struct Ws {
  char *qs, *ks, *vs, *sk;
  float *lse, *dummy, *zeros, *ones, *part;
  size_t sk_bytes;
};

size_t carve(char* base, int B, int M, int Q, Ws* w) {
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += emip_align_up(bytes, 1024); return p; };
  const size_t zeros_n = (size_t)2 * (M > (size_t)B * Q ? M : (size_t)B * Q);
  char* qs = take(match_tc_split_bytes(B, Q, KC));
  char* ks = take(match_tc_split_bytes(B, M, KC));
  char* vs = take(pair_bwd_tc_chn_bytes(B, M));
  float* lse = reinterpret_cast<float*>(take(alf((size_t)B * Q)));
  float* dummy = reinterpret_cast<float*>(take(alf((size_t)B * 2 * Q)));
  float* zeros = reinterpret_cast<float*>(take(alf(zeros_n)));
  float* ones = reinterpret_cast<float*>(take(alf((size_t)B * Q)));
  float* part = reinterpret_cast<float*>(take(alf((size_t)pick_ksplit(B, Q, M) * B * KC * Q)));
  const size_t skb = match_tc_streamk_bytes(B, Q, M);
  char* sk = take(skb);
  if (w) { w->qs = qs; w->ks = ks; w->vs = vs; w->lse = lse; w->dummy = dummy; w->zeros = zeros; w->ones = ones; w->part = part;
           w->sk = sk; w->sk_bytes = skb; }
  return off;
}
}  // namespace

// Shared core: out[b][q][:] (layout `out_layout`, batch stride out_stride_b) = softmax_m(Q K^T / sqrt(128)) V.
// q [B][..] has Q tokens, k / v have M tokens, all 128 channels, all in `layout` (CN = [128][n], NC = [n][128]).
static int attention_tc_core(const float* q_in, const float* m_in, const float* m_out, int layout, float* mem,
                             long long mem_stride_b, int out_layout, float* lse, void* workspace, int B, int M, int Q,
                             cudaStream_t st) {
  Ws w;
  carve(static_cast<char*>(workspace), B, M, Q, &w);
  int rc;
  if ((rc = match_tc_split(q_in, nullptr, w.qs, B, Q, KC, layout, 0, st))) return rc;
  if ((rc = match_tc_split(m_in, nullptr, w.ks, B, M, KC, layout, 0, st))) return rc;
  if ((rc = pair_bwd_tc_split_chn(m_out, nullptr, w.vs, B, M, layout, st))) return rc;
  const size_t zeros_n = (size_t)2 * (M > (size_t)B * Q ? M : (size_t)B * Q);
  EMIP_CUDA(cudaMemsetAsync(w.zeros, 0, zeros_n * sizeof(float), st));
  fill_kernel<<<64, 256, 0, st>>>(w.ones, (size_t)B * Q, -1.0f);
  EMIP_CHECK_LAUNCH("attention_tc (fill)");

  // pass 1: row log-sum-exp
  MatchTcArgs a = {};
  a.x_split = w.qs; a.y_split = w.ks; a.nbx = B; a.nby = B;
  a.v = nullptr; a.v_stride_b = 0; a.grid_w = 64; a.sub_grid = 0;     // analytic-grid mode: only the lse is used
  a.out = w.dummy; a.lse = w.lse;
  a.nb = B; a.nq = Q; a.nk = M; a.y_shift = 0; a.y_mod = B;
  a.s_out = nullptr; a.s_first = 0; a.s_count = 0;
  a.sqrt_c = sqrtf((float)KC);
  a.terms = 3;
  a.sk_ws = w.sk; a.sk_bytes = w.sk_bytes;
  if ((rc = match_tc_fwd(a, st))) return rc;

  // pass 2: out = e^{S - L} V, key tiles split over CTAs when the row tiles alone cannot fill the SMs
  const int ks = pick_ksplit(B, Q, M);
  PairBwdTcArgs t = {};
  t.tok_split = w.qs; t.tok_split_y = w.ks; t.chn_split_y = w.vs;
  t.n_split = B; t.x_base = 0; t.y_base = 0;
  t.l1 = w.lse; t.u = w.zeros; t.u0 = w.ones; t.t = w.zeros; t.t_stride_b = 0;
  t.dx = w.part; t.nb = B; t.nr = Q; t.nc = M; t.dx_layout = out_layout;
  t.sqrt_c = sqrtf((float)KC);
  t.ksplit = ks;
  if ((rc = pair_bwd_tc(t, st))) return rc;
  sum_parts_kernel<<<dim3(128, B), 256, 0, st>>>(w.part, mem, mem_stride_b, ks, B, Q, sqrtf((float)KC));
  EMIP_CHECK_LAUNCH("attention_tc (sum)");
  if (lse != nullptr) EMIP_CUDA(cudaMemcpyAsync(lse, w.lse, (size_t)B * Q * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return EMIP_OK;
}

extern "C" size_t emip_memory_read_tc_workspace(int B, int De, int Do, int M, int Q) {
  if (B < 0 || M <= 0 || Q <= 0 || De != KC || Do != KC) return 0;
  return carve(nullptr, B, M, Q, nullptr);
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
  return carve(nullptr, nb, n, n, nullptr);
}

extern "C" int emip_attention_fwd_tc(const float* q, const float* k, const float* v, float* out, void* workspace,
                                     size_t ws_bytes, int nb, int n, int C, void* stream) {
  if (nb == 0) return EMIP_OK;
  EMIP_CHECK_ARG(q && k && v && out && workspace, "attention_fwd_tc: null pointer");
  EMIP_CHECK_ARG(nb > 0 && n >= 16, "attention_fwd_tc: bad shape nb=%d n=%d (n >= 16)", nb, n);
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
