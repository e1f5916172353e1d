// Token-major layers of the GMFlow FeatureTransformer blocks on the tensor cores (the part that is not
// the attention core): reference model/EMIP_short/motion/gmflow/transformer.py:108-196, TransformerLayer.
//   * the six bias-free nn.Linear layers of a block (q / k / v / merge: 128 -> 128, mlp: 256 -> 1024 -> 128) applied to
//     [L, K] token rows (L = 2B * 1936):  y[l][m] = sum_k act(x[l][k]) w[m][k].  Both operands are K-major as they lie
//     in memory (tokens x channels; nn.Linear stores [out][in]), so the GEMM only needs the elementwise bf16 hi | lo
//     split (gemm_tc.cu, mode 2: three UMMAs per tile, fp32 accumulation in TMEM).  The exact GELU between the two MLP
//     layers (transformer.py:145) is applied inside the operand split of the second layer: the activated [L, 1024]
//     tensor never exists in fp32.
//   * EMIP_LINEAR_W_TRANS: the same contraction against w^T -- the gradient with respect to the input rows (the GMFlow
//     weights are frozen, train.py:340-342, but gradients pass through to the prompt-fusion outputs).
//   * LayerNorm(128) over the channel axis with the residual add of transformer.py:176 folded in, one warp per token
//     row (values stay in registers: one 16-byte load per lane), forward and backward.
#include "common.cuh"
#include "../../include/emip_b200.h"
#include "gemm_tc.cuh"
#include "mlp_fused.cuh"

namespace {
// wt[k][m] = w[m][k]  (the weights are at most 1 MB: 32 x 32 tiles through shared memory)
__global__ void __launch_bounds__(256)
transpose_w_kernel(const float* __restrict__ w, float* __restrict__ wt, int M, int K) {
  __shared__ float t[32][33];
  const int k0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8)
    if (m0 + r < M && k0 + tx < K) t[r][tx] = __ldg(w + (size_t)(m0 + r) * K + k0 + tx);
  __syncthreads();
  for (int r = ty; r < 32; r += 8)
    if (k0 + r < K && m0 + tx < M) wt[(size_t)(k0 + r) * M + m0 + tx] = t[tx][r];
}

size_t wt_bytes(int M, int K) { return emip_align_up((size_t)M * K * sizeof(float), 1024); }
bool shape_ok(int L, int M, int K) { return L >= 0 && M >= 1 && M % 4 == 0 && K >= 16 && K % 4 == 0 && M <= 4096 && K <= 4096; }

// y[l][:] = (res ? res[l][:] : 0) + LN(x[l][:]) * gamma + beta, 128 channels: one warp per row, 8 rows per block
__global__ void __launch_bounds__(256)
ln128_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 const float* __restrict__ res, float* __restrict__ y, int L, float eps) {
  const int lane = threadIdx.x & 31;
  const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + lane), be = __ldg(reinterpret_cast<const float4*>(beta) + lane);
  for (int l = blockIdx.x * 8 + (threadIdx.x >> 5); l < L; l += gridDim.x * 8) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + (size_t)l * 128) + lane);
    const float mean = warp_sum((v.x + v.y) + (v.z + v.w)) * (1.f / 128.f);
    const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
    const float var = warp_sum((a * a + b * b) + (c * c + d * d)) * (1.f / 128.f);
    const float rstd = rsqrtf(var + eps);
    float4 o = make_float4(a * rstd * g.x + be.x, b * rstd * g.y + be.y, c * rstd * g.z + be.z, d * rstd * g.w + be.w);
    if (res != nullptr) {
      const float4 r = __ldg(reinterpret_cast<const float4*>(res + (size_t)l * 128) + lane);
      o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
    }
    reinterpret_cast<float4*>(y + (size_t)l * 128)[lane] = o;
  }
}

// dx[l][:] = rstd (g - mean(g) - xhat mean(g xhat)),  g = dy gamma  (statistics recomputed from x: 512 B per row)
__global__ void __launch_bounds__(256)
ln128_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ dy,
                 float* __restrict__ dx, int L, float eps) {
  const int lane = threadIdx.x & 31;
  const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + lane);
  for (int l = blockIdx.x * 8 + (threadIdx.x >> 5); l < L; l += gridDim.x * 8) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + (size_t)l * 128) + lane);
    const float4 u = __ldg(reinterpret_cast<const float4*>(dy + (size_t)l * 128) + lane);
    const float mean = warp_sum((v.x + v.y) + (v.z + v.w)) * (1.f / 128.f);
    float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
    const float var = warp_sum((a * a + b * b) + (c * c + d * d)) * (1.f / 128.f);
    const float rstd = rsqrtf(var + eps);
    a *= rstd; b *= rstd; c *= rstd; d *= rstd;
    const float ga = u.x * g.x, gb = u.y * g.y, gc = u.z * g.z, gd = u.w * g.w;
    const float m1 = warp_sum((ga + gb) + (gc + gd)) * (1.f / 128.f);
    const float m2 = warp_sum((ga * a + gb * b) + (gc * c + gd * d)) * (1.f / 128.f);
    reinterpret_cast<float4*>(dx + (size_t)l * 128)[lane] =
        make_float4(rstd * (ga - m1 - a * m2), rstd * (gb - m1 - b * m2), rstd * (gc - m1 - c * m2), rstd * (gd - m1 - d * m2));
  }
}
}  // namespace

extern "C" size_t emip_linear_tm_workspace(int L, int M, int K) {
  if (!shape_ok(L, M, K)) return 0;
  const size_t a = gemm_nt_tc_scratch_bytes(1, L, M, K), b = gemm_nt_tc_scratch_bytes(1, L, K, M);
  return emip_align_up(a > b ? a : b, 1024) + wt_bytes(M, K);
}

// y [L][M] = act(x [L][K]) w^T, w [M][K];   EMIP_LINEAR_W_TRANS: y [L][K] = act(x [L][M]) w
extern "C" int emip_linear_tm_fwd(const float* x, const float* w, float* y, void* workspace, size_t ws_bytes, int L, int M, int K,
                                  int flags, void* stream) {
  return emip_linear_tm_fwd_ex(x, nullptr, w, y, workspace, ws_bytes, L, M, K, flags, stream);
}

// y [L][128] = (res ? res : 0) + LayerNorm_128(act(x) w^T) * gamma + beta: merge + norm1 (+ source) and, with
// EMIP_LINEAR_GELU_IN off and the rows pre-activated by emip_mlp_ln_tm_fwd, mlp[2] + norm2 + source
extern "C" int emip_linear_ln_tm_fwd(const float* x, const float* w, const float* gamma, const float* beta, const float* res,
                                     float* y, void* workspace, size_t ws_bytes, int L, int M, int K, float eps, int flags,
                                     void* stream) {
  if (L == 0) return EMIP_OK;
  EMIP_CHECK_ARG(x && w && y && gamma && beta && workspace, "linear_ln_tm_fwd: null pointer");
  EMIP_CHECK_ARG((flags & ~EMIP_LINEAR_GELU_IN) == 0, "linear_ln_tm_fwd: unknown flag");
  if (M != 128 || !shape_ok(L, M, K)) { emip_set_error("linear_ln_tm_fwd: unsupported shape L=%d M=%d K=%d (M must be 128)", L, M, K); return EMIP_ENOSYS; }
  if (ws_bytes < emip_linear_tm_workspace(L, M, K)) { emip_set_error("linear_ln_tm_fwd: workspace too small"); return EMIP_ENOMEM; }
  GemmNT t = {};
  t.B = 1; t.M = L; t.K = M; t.N = K;
  t.a = x; t.lda = K;
  t.bm = w; t.ldb = K;
  t.c = y; t.ldc = M;
  t.a_act = (flags & EMIP_LINEAR_GELU_IN) ? 1 : 0;
  t.ln_gamma = gamma; t.ln_beta = beta; t.ln_eps = eps; t.c_res = res;
  if (!gemm_nt_tc_supported(t)) { emip_set_error("linear_ln_tm_fwd: output must be 16-byte aligned"); return EMIP_EINVAL; }
  return gemm_nt_tc(t, workspace, ws_bytes - wt_bytes(M, K), (cudaStream_t)stream, 1);
}

// EMIP_LINEAR_GELU_BWD_IN: the rows are x * GELU'(aux) (aux laid out like x) -- the gradient entering the first MLP layer
extern "C" int emip_linear_tm_fwd_ex(const float* x, const float* aux, const float* w, float* y, void* workspace, size_t ws_bytes,
                                     int L, int M, int K, int flags, void* stream) {
  if (L == 0) return EMIP_OK;
  EMIP_CHECK_ARG(x && w && y && workspace, "linear_tm_fwd: null pointer");
  EMIP_CHECK_ARG((flags & ~(EMIP_LINEAR_GELU_IN | EMIP_LINEAR_W_TRANS | EMIP_LINEAR_GELU_BWD_IN)) == 0, "linear_tm_fwd: unknown flag");
  EMIP_CHECK_ARG(!(flags & EMIP_LINEAR_GELU_BWD_IN) || (aux != nullptr && !(flags & EMIP_LINEAR_GELU_IN)),
                 "linear_tm_fwd: EMIP_LINEAR_GELU_BWD_IN needs aux and excludes EMIP_LINEAR_GELU_IN");
  if (!shape_ok(L, M, K)) { emip_set_error("linear_tm_fwd: unsupported shape L=%d M=%d K=%d", L, M, K); return EMIP_ENOSYS; }
  if (ws_bytes < emip_linear_tm_workspace(L, M, K)) { emip_set_error("linear_tm_fwd: workspace too small"); return EMIP_ENOMEM; }
  EMIP_CHECK_ARG(reinterpret_cast<uintptr_t>(workspace) % 1024 == 0, "linear_tm_fwd: workspace must be 1024-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t scratch_bytes = ws_bytes - wt_bytes(M, K);
  const bool trans = (flags & EMIP_LINEAR_W_TRANS) != 0;
  const float* wb = w;
  if (trans) {
    float* wt = reinterpret_cast<float*>(static_cast<char*>(workspace) + scratch_bytes);
    transpose_w_kernel<<<dim3((K + 31) / 32, (M + 31) / 32), 256, 0, st>>>(w, wt, M, K);
    EMIP_CHECK_LAUNCH("linear_tm_fwd (transpose)");
    wb = wt;
  }
  const int out = trans ? K : M, in = trans ? M : K;
  GemmNT t = {};
  t.B = 1; t.M = L; t.K = out; t.N = in;
  t.a = x; t.a_stride_b = 0; t.lda = in;
  t.bm = wb; t.b_stride_b = 0; t.ldb = in;
  t.c = y; t.c_stride_b = 0; t.ldc = out;
  t.a_act = (flags & EMIP_LINEAR_GELU_IN) ? 1 : (flags & EMIP_LINEAR_GELU_BWD_IN) ? 2 : 0;
  t.a_aux = aux;
  if (!gemm_nt_tc_supported(t)) { emip_set_error("linear_tm_fwd: output must be 16-byte aligned"); return EMIP_EINVAL; }
  return gemm_nt_tc(t, workspace, scratch_bytes, st, 1);
}

// y [L][M] = x [L][K] w^T + bias [M]: nn.Linear with bias on token rows (FeatureFlowAttention.q_proj / k_proj on the
// transformer's token-major output, transformer.py:523-524); forward only
extern "C" int emip_linear_tm_bias_fwd(const float* x, const float* w, const float* bias, float* y, void* workspace, size_t ws_bytes,
                                       int L, int M, int K, void* stream) {
  if (L == 0) return EMIP_OK;
  EMIP_CHECK_ARG(x && w && y && workspace, "linear_tm_bias_fwd: null pointer");
  if (!shape_ok(L, M, K)) { emip_set_error("linear_tm_bias_fwd: unsupported shape L=%d M=%d K=%d", L, M, K); return EMIP_ENOSYS; }
  if (ws_bytes < emip_linear_tm_workspace(L, M, K)) { emip_set_error("linear_tm_bias_fwd: workspace too small"); return EMIP_ENOMEM; }
  EMIP_CHECK_ARG(reinterpret_cast<uintptr_t>(workspace) % 1024 == 0, "linear_tm_bias_fwd: workspace must be 1024-byte aligned");
  GemmNT t = {};
  t.B = 1; t.M = L; t.K = M; t.N = K;
  t.a = x; t.lda = K;
  t.bm = w; t.ldb = K;
  t.c = y; t.ldc = M;
  t.c_bias = bias;
  if (!gemm_nt_tc_supported(t)) { emip_set_error("linear_tm_bias_fwd: output must be 16-byte aligned"); return EMIP_EINVAL; }
  return gemm_nt_tc(t, workspace, ws_bytes - wt_bytes(M, K), (cudaStream_t)stream, 1);
}

// y_i [L][M] = x [L][K] w_i^T for n weights of one shape: the rows are split once (q / k / v of a self-attention layer read the
// same tokens, k / v of a cross-attention layer too)
extern "C" size_t emip_linear_tm_multi_workspace(int L, int M, int K) {
  if (!shape_ok(L, M, K) || K % 64 != 0) return 0;
  return emip_align_up((size_t)L * K * 2 * 2, 1024) + gemm_nt_tc_scratch_bytes_presplit(M, K);
}

extern "C" int emip_linear_tm_multi_fwd(const float* x, const float* const* w, float* const* y, int n, void* workspace, size_t ws_bytes,
                                        int L, int M, int K, void* stream) {
  if (L == 0 || n == 0) return EMIP_OK;
  EMIP_CHECK_ARG(x && w && y && workspace && n > 0, "linear_tm_multi_fwd: null pointer");
  if (!shape_ok(L, M, K) || K % 64 != 0) { emip_set_error("linear_tm_multi_fwd: unsupported shape L=%d M=%d K=%d (K %% 64 == 0)", L, M, K); return EMIP_ENOSYS; }
  if (ws_bytes < emip_linear_tm_multi_workspace(L, M, K)) { emip_set_error("linear_tm_multi_fwd: workspace too small"); return EMIP_ENOMEM; }
  EMIP_CHECK_ARG(reinterpret_cast<uintptr_t>(workspace) % 1024 == 0, "linear_tm_multi_fwd: workspace must be 1024-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* hi = static_cast<__nv_bfloat16*>(workspace);
  __nv_bfloat16* lo = hi + (size_t)L * K;
  const size_t split_bytes = emip_align_up((size_t)L * K * 2 * 2, 1024);
  int rc;
  if ((rc = gemm_tc_split_rows(x, L, K, hi, lo, st))) return rc;
  for (int i = 0; i < n; ++i) {
    EMIP_CHECK_ARG(w[i] && y[i], "linear_tm_multi_fwd: null weight / output pointer");
    GemmNT t = {};
    t.B = 1; t.M = L; t.K = M; t.N = K;
    t.a_hi_pre = hi; t.a_lo_pre = lo;
    t.bm = w[i]; t.ldb = K;
    t.c = y[i]; t.ldc = M;
    if (!gemm_nt_tc_supported(t)) { emip_set_error("linear_tm_multi_fwd: output must be 16-byte aligned"); return EMIP_EINVAL; }
    if ((rc = gemm_nt_tc(t, static_cast<char*>(workspace) + split_bytes, ws_bytes - split_bytes, st, 1))) return rc;
  }
  return EMIP_OK;
}

namespace {
size_t hid_bytes(int L, int Hd) { return emip_align_up((size_t)L * Hd * 2 * 2, 1024); }      // bf16 hi + lo
bool mlp_shape_ok(int L, int K1, int Hd, int M) {
  return L >= 0 && K1 >= 16 && K1 % 4 == 0 && K1 <= 4096 && Hd >= 64 && Hd % 64 == 0 && Hd <= 8192 && M >= 1 && M % 4 == 0 && M <= 4096;
}
}  // namespace

extern "C" size_t emip_mlp_tm_workspace(int L, int K1, int Hd, int M) {
  if (!mlp_shape_ok(L, K1, Hd, M)) return 0;
  const size_t s1 = gemm_nt_tc_scratch_bytes(1, L, Hd, K1), s2 = gemm_nt_tc_scratch_bytes_presplit(M, Hd);
  return hid_bytes(L, Hd) + emip_align_up(s1 > s2 ? s1 : s2, 1024);
}

// y [L][M] = GELU(x [L][K1] w1^T) w2^T,  w1 [Hd][K1], w2 [M][Hd]: the epilogue of the first GEMM writes the activated
// hidden rows as the bf16 hi | lo A operand of the second one (no fp32 hidden tensor, no separate split pass)
extern "C" int emip_mlp_tm_fwd(const float* x, const float* w1, const float* w2, float* y, void* workspace, size_t ws_bytes, int L,
                               int K1, int Hd, int M, void* stream) {
  return emip_mlp_ln_tm_fwd(x, w1, w2, nullptr, nullptr, nullptr, y, workspace, ws_bytes, L, K1, Hd, M, 0.f, stream);
}

// gamma != NULL (M == 128): y = (res ? res : 0) + LayerNorm_128(GELU(x w1^T) w2^T) * gamma + beta -- mlp + norm2 + source
extern "C" int emip_mlp_ln_tm_fwd(const float* x, const float* w1, const float* w2, const float* gamma, const float* beta,
                                  const float* res, float* y, void* workspace, size_t ws_bytes, int L, int K1, int Hd, int M,
                                  float eps, void* stream) {
  if (L == 0) return EMIP_OK;
  EMIP_CHECK_ARG(gamma == nullptr || (beta != nullptr && M == 128), "mlp_ln_tm_fwd: the LayerNorm epilogue needs beta and M = 128");
  EMIP_CHECK_ARG(x && w1 && w2 && y && workspace, "mlp_tm_fwd: null pointer");
  if (!mlp_shape_ok(L, K1, Hd, M)) { emip_set_error("mlp_tm_fwd: unsupported shape L=%d K1=%d H=%d M=%d", L, K1, Hd, M); return EMIP_ENOSYS; }
  if (ws_bytes < emip_mlp_tm_workspace(L, K1, Hd, M)) { emip_set_error("mlp_tm_fwd: workspace too small"); return EMIP_ENOMEM; }
  EMIP_CHECK_ARG(reinterpret_cast<uintptr_t>(workspace) % 1024 == 0, "mlp_tm_fwd: workspace must be 1024-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  // the transformer's shape with its LayerNorm: one kernel, hidden rows kept on the SM (mlp_fused.cu); the workspace then only holds
  // the operand splits of x, w1, w2 (always smaller than the hidden tensor it is sized for, except for a handful of rows)
  if (gamma != nullptr && K1 == 256 && M == 128 && Hd % 128 == 0 && !gemm_tc_debug_two_launch_mlp()) {
    const size_t xb = emip_align_up((size_t)L * K1 * 2 * 2, 1024), w1b = emip_align_up(gemm_tc_split_b_bytes(Hd, K1), 1024),
                 w2b = emip_align_up(gemm_tc_split_b_bytes(M, Hd), 1024);
    MlpFusedArgs f = {};
    char* base = static_cast<char*>(workspace);
    f.x_hi = base; f.x_lo = base + (size_t)L * K1 * 2; f.ldx = K1;
    f.w1 = base + xb; f.w2 = base + xb + w1b; f.L = L; f.hid = Hd;
    f.gamma = gamma; f.beta = beta; f.eps = eps; f.res = res; f.ldr = M; f.y = y; f.ldy = M;
    if (xb + w1b + w2b <= ws_bytes && mlp_fused_supported(f)) {
      int rc;
      if ((rc = gemm_tc_split_rows(x, L, K1, base, base + (size_t)L * K1 * 2, st))) return rc;
      if ((rc = gemm_tc_split_b(w1, K1, Hd, K1, base + xb, st))) return rc;
      if ((rc = gemm_tc_split_b(w2, Hd, M, Hd, base + xb + w1b, st))) return rc;
      return mlp_fused_tc(f, st);
    }
  }
  __nv_bfloat16* h_hi = static_cast<__nv_bfloat16*>(workspace);
  __nv_bfloat16* h_lo = h_hi + (size_t)L * Hd;
  void* scratch = static_cast<char*>(workspace) + hid_bytes(L, Hd);
  const size_t scratch_bytes = ws_bytes - hid_bytes(L, Hd);
  int rc;
  GemmNT t = {};
  t.B = 1; t.M = L; t.K = Hd; t.N = K1;
  t.a = x; t.lda = K1;
  t.bm = w1; t.ldb = K1;
  t.ldc = Hd;
  t.c_hi = h_hi; t.c_lo = h_lo; t.ldc_split = Hd; t.c_act = 1;
  if (!gemm_nt_tc_supported(t)) { emip_set_error("mlp_tm_fwd: unsupported first layer"); return EMIP_ENOSYS; }
  if ((rc = gemm_nt_tc(t, scratch, scratch_bytes, st, 1))) return rc;
  GemmNT u = {};
  u.B = 1; u.M = L; u.K = M; u.N = Hd;
  u.a_hi_pre = h_hi; u.a_lo_pre = h_lo;
  u.bm = w2; u.ldb = Hd;
  u.c = y; u.ldc = M;
  u.ln_gamma = gamma; u.ln_beta = beta; u.ln_eps = eps; u.c_res = gamma ? res : nullptr;
  if (!gemm_nt_tc_supported(u)) { emip_set_error("mlp_tm_fwd: output must be 16-byte aligned"); return EMIP_EINVAL; }
  return gemm_nt_tc(u, scratch, scratch_bytes, st, 1);
}

extern "C" int emip_layernorm_tm_fwd(const float* x, const float* gamma, const float* beta, const float* res, float* y, int L,
                                     int C, float eps, void* stream) {
  if (L == 0) return EMIP_OK;
  EMIP_CHECK_ARG(x && gamma && beta && y && L > 0, "layernorm_tm_fwd: null pointer");
  if (C != 128) { emip_set_error("layernorm_tm_fwd: C=%d (only the FeatureTransformer width 128 is built)", C); return EMIP_ENOSYS; }
  EMIP_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(res) |
                   reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15) == 0,
                 "layernorm_tm_fwd: pointers must be 16-byte aligned");
  const int blocks = (L + 7) / 8, cap = emip_num_sms() * 8;
  ln128_fwd_kernel<<<blocks < cap ? blocks : cap, 256, 0, (cudaStream_t)stream>>>(x, gamma, beta, res, y, L, eps);
  EMIP_CHECK_LAUNCH("layernorm_tm_fwd");
  return EMIP_OK;
}

extern "C" int emip_layernorm_tm_bwd(const float* x, const float* gamma, const float* dy, float* dx, int L, int C, float eps,
                                     void* stream) {
  if (L == 0) return EMIP_OK;
  EMIP_CHECK_ARG(x && gamma && dy && dx && L > 0, "layernorm_tm_bwd: null pointer");
  if (C != 128) { emip_set_error("layernorm_tm_bwd: C=%d (only the FeatureTransformer width 128 is built)", C); return EMIP_ENOSYS; }
  EMIP_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx) |
                   reinterpret_cast<uintptr_t>(gamma)) & 15) == 0, "layernorm_tm_bwd: pointers must be 16-byte aligned");
  const int blocks = (L + 7) / 8, cap = emip_num_sms() * 8;
  ln128_bwd_kernel<<<blocks < cap ? blocks : cap, 256, 0, (cudaStream_t)stream>>>(x, gamma, dy, dx, L, eps);
  EMIP_CHECK_LAUNCH("layernorm_tm_bwd");
  return EMIP_OK;
}
