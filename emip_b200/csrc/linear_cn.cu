// Channel-major linear layer on the tensor cores: y[b][m][n] = sum_k w[m][k] x[b][k][n] + bias[m]  (x, y [B][C][N]).
// a2: the two projections of FeatureFlowAttention (reference model/EMIP_short/motion/gmflow/transformer.py:523-524,
// q = q_proj(feature0), k = k_proj(q)) applied to the feature map as it lies in memory ([B,C,H,W]), so that neither the
// input nor the projected maps are ever permuted to token-major fp32; forward and backward are split-bf16 tcgen05 GEMMs
// (gemm_tc.cu: the activations are read as an MN-major operand, the weight gradient is a split-K GEMM over the pixels).
#include "common.cuh"
#include "../../include/emip_b200.h"
#include "gemm_tc.cuh"

namespace {
// part[b][m] = sum_n dy[b][m][n]  (N % 4 == 0; the batch is summed afterwards in a fixed order)
__global__ void __launch_bounds__(128)
row_sum_kernel(const float* __restrict__ dy, float* __restrict__ part, int M, int N) {
  __shared__ float red[4];
  const int m = blockIdx.x, b = blockIdx.y;
  const float4* src = reinterpret_cast<const float4*>(dy + ((size_t)b * M + m) * N);
  float s = 0.f;
  for (int i = threadIdx.x; i < (N >> 2); i += blockDim.x) {
    const float4 v = __ldg(src + i);
    s += (v.x + v.y) + (v.z + v.w);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) part[(size_t)b * M + m] = (red[0] + red[1]) + (red[2] + red[3]);
}
int nt_splits(int B, int M, int K) {
  const int tiles = B * ((M + 127) / 128) * ((K + 127) / 128);
  int s = emip_num_sms() / tiles;
  return s < 1 ? 1 : (s > 8 ? 8 : s);
}
size_t part_bytes(int B, int M, int K) { return emip_align_up((size_t)B * nt_splits(B, M, K) * M * K * sizeof(float), 1024); }
bool shape_ok(int B, int M, int K, int N) { return B >= 0 && M >= 1 && K >= 1 && K <= 1024 && M <= 1024 && N >= 16 && N % 4 == 0; }
}  // namespace

extern "C" size_t emip_linear_cn_workspace(int B, int M, int K, int N) {
  if (!shape_ok(B, M, K, N)) return 0;
  size_t s = gemm_nn_tc_scratch_bytes(B, M, K, N, false), t = gemm_nn_tc_scratch_bytes(B, K, M, N, false);
  if (t > s) s = t;
  t = gemm_nt_tc_scratch_bytes(B, M, K, N);
  if (t > s) s = t;
  return emip_align_up(s, 1024) + part_bytes(B, M, K);
}

extern "C" int emip_linear_cn_fwd(const float* x, const float* w, const float* bias, float* y, void* workspace, size_t ws_bytes,
                                  int B, int M, int K, int N, void* stream) {
  if (B == 0) return EMIP_OK;
  EMIP_CHECK_ARG(x && w && y && workspace, "linear_cn_fwd: null pointer");
  if (!shape_ok(B, M, K, N)) { emip_set_error("linear_cn_fwd: unsupported shape M=%d K=%d N=%d", M, K, N); return EMIP_ENOSYS; }
  if (ws_bytes < emip_linear_cn_workspace(B, M, K, N)) { emip_set_error("linear_cn_fwd: workspace too small"); return EMIP_ENOMEM; }
  GemmNN a = {};
  a.B = B; a.M = M; a.K = K; a.N = N;
  a.w = w; a.w_stride_b = 0; a.ldw = K; a.w_trans = 0;
  a.x = x; a.x_stride_b = (long long)K * N; a.ldx = N;
  a.y = y; a.y_stride_b = (long long)M * N; a.ldy = N;
  a.bias = bias;
  if (!gemm_nn_tc_supported(a)) { emip_set_error("linear_cn_fwd: output must be 16-byte aligned"); return EMIP_EINVAL; }
  return gemm_nn_tc(a, workspace, ws_bytes - part_bytes(B, M, K), (cudaStream_t)stream);
}

// dx [B][K][N] = w^T dy;  dw [M][K] = sum_b dy[b] x[b]^T (NULL: skipped);  db [M] = sum dy (NULL: skipped)
extern "C" int emip_linear_cn_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, float* db, void* workspace,
                                  size_t ws_bytes, int B, int M, int K, int N, void* stream) {
  if (B == 0) return EMIP_OK;
  EMIP_CHECK_ARG(x && w && dy && workspace, "linear_cn_bwd: null pointer");
  if (!shape_ok(B, M, K, N)) { emip_set_error("linear_cn_bwd: unsupported shape M=%d K=%d N=%d", M, K, N); return EMIP_ENOSYS; }
  if (ws_bytes < emip_linear_cn_workspace(B, M, K, N)) { emip_set_error("linear_cn_bwd: workspace too small"); return EMIP_ENOMEM; }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t scratch_bytes = ws_bytes - part_bytes(B, M, K);
  int rc;
  if (dx != nullptr) {
    GemmNN a = {};
    a.B = B; a.M = K; a.K = M; a.N = N;
    a.w = w; a.w_stride_b = 0; a.ldw = K; a.w_trans = 1;                 // W^T
    a.x = dy; a.x_stride_b = (long long)M * N; a.ldx = N;
    a.y = dx; a.y_stride_b = (long long)K * N; a.ldy = N;
    if (!gemm_nn_tc_supported(a)) { emip_set_error("linear_cn_bwd: dx must be 16-byte aligned"); return EMIP_EINVAL; }
    if ((rc = gemm_nn_tc(a, workspace, scratch_bytes, st))) return rc;
  }
  if (dw != nullptr) {
    float* part = reinterpret_cast<float*>(static_cast<char*>(workspace) + scratch_bytes);
    const int ns = nt_splits(B, M, K);
    GemmNT t = {};
    t.B = B; t.M = M; t.K = K; t.N = N;
    t.a = dy; t.a_stride_b = (long long)M * N; t.lda = N;
    t.bm = x; t.b_stride_b = (long long)K * N; t.ldb = N;
    t.c = part; t.c_stride_b = (long long)M * K; t.ldc = K;
    if (!gemm_nt_tc_supported(t)) { emip_set_error("linear_cn_bwd: unsupported weight-gradient shape"); return EMIP_ENOSYS; }
    if ((rc = gemm_nt_tc(t, workspace, scratch_bytes, st, ns))) return rc;
    if ((rc = reduce_batch(part, (long long)M * K, dw, B * ns, (long long)M * K, 0, st))) return rc;
  }
  if (db != nullptr) {
    float* part = reinterpret_cast<float*>(static_cast<char*>(workspace) + scratch_bytes);     // free again: dw is reduced
    row_sum_kernel<<<dim3(M, B), 128, 0, st>>>(dy, part, M, N);
    EMIP_CHECK_LAUNCH("linear_cn_bwd (bias)");
    if ((rc = reduce_batch(part, M, db, B, M, 0, st))) return rc;
  }
  return EMIP_OK;
}
