/* emip_b200 -- C ABI of the B200-native EMIP motion-stream hot path.
 *
 * The reference (zhangxin06/EMIP) has no FFI layer: its hot path is four Python
 * functions/modules that call ATen.  Each entry point below replaces one of
 * those call sites; the citation names the reference interface it stands in for
 * (paths relative to the reference root).  The Python host side
 * (emip_b200/*.py) binds these with ctypes; INTEGRATION.md shows the binding a
 * reference maintainer would add.
 *
 * Conventions (all entry points):
 *  - plain C: device pointers + sizes, no torch types.  Pointers are CUDA device
 *    pointers on the current device; fp32 unless stated otherwise.
 *  - the CALLER owns every buffer (outputs, saved-for-backward, workspace); the
 *    library never allocates device memory, never synchronises the stream and
 *    never throws.  `stream` is a cudaStream_t passed as void*.
 *  - returns 0 on success, a negative errno-style code for argument errors
 *    (-22 EINVAL, -38 ENOSYS, -12 ENOMEM = workspace too small) or a positive
 *    cudaError_t for launch failures; emip_last_error() holds the message
 *    (thread-local).
 *  - re-entrant: forward is called on the Python thread, backward on autograd
 *    engine threads.  Compute entry points keep NO state between calls.  The
 *    complete list of process-wide mutable state in the library:
 *      (1) one mutex-guarded cache in csrc/abi.cu: the SM count per device and the
 *          set of (kernel, device) pairs whose dynamic-shared-memory limit has been
 *          raised (cudaFuncSetAttribute is per device) -- one process may drive
 *          several GPUs; the pointers must belong to the CURRENT device and the
 *          stream to that device (the Python host wraps calls in
 *          torch.cuda.device(tensor.device));
 *      (2) the address of the driver's cuTensorMapEncodeTiled, resolved once;
 *      (3) thread-local: the last error string, and flow_warp's two most recent
 *          tensor maps (keyed by pointer / shape / strides; never shared);
 *      (4) diagnostic switches for the profiling scripts under tools/ --
 *          emip_match_tc_set_profile_buffer, emip_match_tc_set_variant,
 *          emip_attn_tc_set_profile_buffer, emip_gemm_tc_set_profile_buffer,
 *          emip_debug_flow_warp_staged_profile, emip_debug_gemm_wide_tiles, and the launch policy
 *          emip_set_programmatic_launch (default off) --
 *          NOT thread-safe, never touched by the host package, default off.
 *    No entry point allocates or frees memory (cudaMalloc / cudaHostAlloc), so all
 *    of them may be captured into CUDA graphs once (1) is warm (first call).
 *  - there is NO CPU fallback: on a machine without an sm_100 device every
 *    compute entry point fails.
 */
#ifndef EMIP_B200_H
#define EMIP_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EMIP_ABI_VERSION 4

#define EMIP_PAD_BORDER 0
#define EMIP_PAD_ZEROS 1

/* flags */
#define EMIP_FLAG_EXACT_FP32 1 /* force the exact-fp32 CUDA-core path (no bf16 hi/lo tensor-core split) */
#define EMIP_FLAG_REUSE_WORKSPACE 2 /* global_matching_fwd: the workspace still holds the operand split and pixel
                                       grid of the previous call on the SAME f0/f1; skip those pre-passes (used to
                                       time / profile the fused kernel alone) */

#define EMIP_FLAG_BF16 4            /* bf16 inference mode: single-pass bf16 operands (no hi/lo split), fp32 accumulate and
                                       fp32 softmax; results within the 2e-2 tolerance instead of 1e-3 */
#define EMIP_FLAG_CHANNEL_MAJOR 8  /* flow_attn_fwd / _bwd: q, k (and dq, dk) are channel-major [B,C,N] instead of [B,N,C] */
#define EMIP_FLAG_TOKEN_MAJOR 16   /* global_matching_fwd: f0, f1 are token-major [B,H*W,C] -- the FeatureTransformer's own
                                       output layout (transformer.py:476, before the permute of :479-480) */
#define EMIP_FLAG_PRESPLIT 128      /* global_matching_fwd: f0 is not fp32 features but the bf16 operand [2B][H*W][hi 128 | lo 128] of
                                     * BOTH frames (frame-1 maps, then frame-2 maps) as emip_feature_transformer_fwd_ex writes it
                                     * (out_split); f1 is ignored and no operand-split pass runs.  Tensor-core path only. */
#define EMIP_FLAG_SCHED_STREAMK 32  /* global_matching_fwd / flow_attn_fwd: force the stream-K work schedule (default: chosen by */
#define EMIP_FLAG_SCHED_ITEMS 64    /* batch size) / force grid-strided whole items -- results are identical either way */
#define EMIP_LAYOUT_TOKEN_MAJOR 0  /* [B][H*W][C] */
#define EMIP_LAYOUT_CHANNEL_MAJOR 1 /* [B][C][H*W]  (NCHW feature maps) */

/* ---- plumbing ------------------------------------------------------------ */
const char* emip_last_error(void);
int emip_abi_version(void);
/* 0 when the current CUDA device is sm_100 (B200); error otherwise. */
int emip_device_check(void);

/* Diagnostics for the tcgen05 kernel behind a1/a2: when set to a device buffer of (SM count) x 8 uint64, every
 * following launch stores per-CTA wait-cycle counters
 * [producer: q_empty, k_empty | issuer: s_empty, k_full, q_full, total | softmax warp 2: s_full, total].
 * NULL (the default) switches it off.  Not thread-safe; for profiling scripts only. */
void emip_match_tc_set_profile_buffer(unsigned long long* dev_buf);
/* Diagnostics: force 8 or 16 softmax warps per CTA in the same kernel (0 = default choice). */
void emip_match_tc_set_variant(int softmax_warps);
/* Diagnostics: role wait-cycle profile of gemm_tc_kernel, device pointer to [SM count][8] uint64 or NULL (tools/gemm_roles.py). */
void emip_gemm_tc_set_profile_buffer(unsigned long long* dev_buf);
/* Diagnostics / tuning, a bit set.  Bit 0: gemm_tc tile width for outputs of >= 512 columns: 0 = 128-column tiles, four TMEM
 * accumulators (default), 1 = 256-column tiles.  Bit 1: the epilogue warps of gemm_tc_kernel skip their work (no math, no stores:
 * the OUTPUT IS GARBAGE) so that tools/gemm_floor.py can time the TMA + MMA pipeline alone.  Bit 2: emip_feature_transformer_fwd runs
 * the feed-forward network as two GEMM launches (hidden rows through HBM) instead of the fused kernel -- for A / B timing and the
 * parity test of the fused kernel.  Bit 3: the producer warp skips the operand loads as well (UMMA issue rate alone; garbage).
 * 0 restores the defaults. */
void emip_debug_gemm_wide_tiles(int v);
/* Launch policy, process-wide, default 0; returns the previous value.  1: the persistent tensor-core kernels (gemm_tc / attention /
 * fused feed-forward) are launched with programmatic stream serialisation -- a kernel's prologue (barrier init, TMEM allocation)
 * may start on an SM as soon as the previous kernel's CTA there has exited, and waits (griddepcontrol.wait) for that kernel to
 * complete before it touches memory.  Worth ~1 % of the chain step when ONE stream owns the GPU (bench.py c3; captured into a CUDA
 * graph the setting is frozen into the graph's edges); leave it off when several streams share the GPU: an early-launched CTA
 * holds an SM that another stream's kernel could use (c4 with four clip streams per GPU: -28 %).  Not thread-safe: set it before
 * launching / capturing, e.g. emip_b200.chain.GraphedChain(pdl=True) sets it around its capture only. */
int emip_set_programmatic_launch(int on);
/* Diagnostics: wait-cycle profile of the staged flow_warp kernel, device pointer to [grid][8] int64 or NULL. */
void emip_debug_flow_warp_staged_profile(long long* buf);

/* ---- a3: flow_warp ------------------------------------------------------- */
/* Replaces loss/warp_utils.py:83-93 flow_warp(x, flow12, pad, mode='bilinear')
 * (mesh_grid :7-13 + norm_grid :16-23 + F.grid_sample(align_corners=True)).
 *   x    [B,C,H,W] contiguous        image to sample
 *   flow element (b,ch,y,x) at flow[b*flow_stride_b + ch*flow_stride_c + y*W + x]
 *        (ch 0 = dx, 1 = dy; strides in elements) -- covers the non-contiguous
 *        flow[:, :2] / flow[:, 2:] slices of loss/loss_flow.py:90-91
 *   out  [B,C,H,W] contiguous
 *   pad_mode EMIP_PAD_BORDER ('border', the photometric loss) or EMIP_PAD_ZEROS. */
int emip_flow_warp_fwd(const float* x, const float* flow, float* out, int B, int C, int H, int W,
                       long long flow_stride_b, long long flow_stride_c, int pad_mode, void* stream);
/* As emip_flow_warp_fwd / _bwd with an explicit kernel choice and CALLER-OWNED kernel-choice feedback (the library keeps
 * none).  For C = 3, border padding, no image gradient there are two bit-identical kernels: a TMA-staged one (taps from
 * shared memory: ~15 % faster on smooth flow such as the model's convex-upsampled output) and a direct-gather one (several
 * times faster than the staged kernel when most 32x32 tiles have a flow range > ~14 px, e.g. iid noise).
 *   kernel     EMIP_WARP_KERNEL_AUTO (staged when shape / alignment allow, else direct), _DIRECT, _STAGED (ENOSYS if not
 *              covered); bits 8.. = persistent CTAs per SM (0 = default 2; tools only)
 *   dev_stats  NULL or 16 bytes of zeroed device memory owned by the caller, not shared by launches in flight at once
 *   host_stats NULL or 8 bytes of device-accessible pinned host memory: when both are given and the staged kernel runs, its
 *              last CTA writes host_stats[1] = per-mille of tiles that fell back to global gathers, then host_stats[0] = seq
 *              (and re-zeroes dev_stats).  emip_b200/warp.py keeps such buffers per device and switches kernels on it. */
#define EMIP_WARP_KERNEL_AUTO 0
#define EMIP_WARP_KERNEL_DIRECT 1
#define EMIP_WARP_KERNEL_STAGED 2
int emip_flow_warp_fwd_ex(const float* x, const float* flow, float* out, int B, int C, int H, int W, long long flow_stride_b,
                          long long flow_stride_c, int pad_mode, int kernel, unsigned* dev_stats, unsigned* host_stats,
                          unsigned seq, void* stream);
int emip_flow_warp_bwd_ex(const float* x, const float* flow, const float* dout, float* dflow, float* dx, int B, int C, int H,
                          int W, long long flow_stride_b, long long flow_stride_c, int pad_mode, int kernel, unsigned* dev_stats,
                          unsigned* host_stats, unsigned seq, void* stream);
/* Backward of emip_flow_warp_fwd.  dflow [B,2,H,W] contiguous is overwritten.  dx may be
 * NULL (images need no gradient in the photometric loss); otherwise it must be
 * zero-filled [B,C,H,W] and receives the scatter-added image gradient. */
int emip_flow_warp_bwd(const float* x, const float* flow, const float* dout, float* dflow, float* dx,
                       int B, int C, int H, int W, long long flow_stride_b, long long flow_stride_c,
                       int pad_mode, void* stream);

/* ---- a1: GMFlow global matching ------------------------------------------ */
/* Replaces model/EMIP_short/motion/gmflow/matching.py:8-41
 *   global_correlation_softmax(feature0, feature1, pred_bidir_flow) -> (flow, prob, corr)
 * S[b,i,j] = f0[b,:,i].f1[b,:,j]/sqrt(C); flow_fw = softmax_j(S) g - g; flow_bw likewise on S^T.
 *   f0, f1  [B,C,H,W] contiguous (C must be 128)
 *   flow    [nd*B,2,H,W] out, nd = bidir ? 2 : 1 (first B forward, last B backward; matching.py:29,39)
 *   corr    NULL, or the scaled scores (matching.py:16-20).  Memory layout: bidir -> [B,HW(j),HW(i)], i.e.
 *           corr[b,k,y,x] contiguous; uni-directional -> [B,HW(i),HW(j)] (the reference's own layout, of which
 *           its `corr` is a permuted view).  `prob` (matching.py:34) is never produced: no caller reads it.
 *   lse     NULL, or [nd*B,HW] row log-sum-exp, required by the backward
 *   workspace  >= emip_global_matching_workspace(B,C,H,W) bytes, 1024-byte aligned
 *   flags   0 = tensor-core path (bf16 hi/lo split, fp32 accumulate); EMIP_FLAG_EXACT_FP32 = CUDA-core fp32 */
size_t emip_global_matching_workspace(int B, int C, int H, int W);
int emip_global_matching_fwd(const float* f0, const float* f1, float* flow, float* corr, float* lse,
                             void* workspace, size_t ws_bytes, int B, int C, int H, int W, int bidir, int flags,
                             void* stream);
/* Backward of the above (what autograd derives for matching.py:16-39).  dflow [nd*B,2,H,W] and/or dcorr (same
 * layout as corr) may be NULL; flow and lse are the forward outputs.  df0, df1 [B,C,H,W] are overwritten.
 * flags: 0 = tensor-core path (S recomputed and dX accumulated with bf16 hi/lo splits, H*W % 8 == 0),
 *        EMIP_FLAG_EXACT_FP32 = CUDA-core fp32. */
int emip_global_matching_bwd(const float* f0, const float* f1, const float* flow, const float* lse,
                             const float* dflow, const float* dcorr, float* df0, float* df1, void* workspace,
                             size_t ws_bytes, int B, int C, int H, int W, int bidir, int flags, void* stream);

/* ---- a2: flow-propagation attention -------------------------------------- */
/* Replaces the attention core of FeatureFlowAttention.forward,
 * model/EMIP_short/motion/gmflow/transformer.py:526-532: out = softmax(q k^T / sqrt(C)) v.
 *   q, k  [B,N,C] token-major ([B,C,N] with EMIP_FLAG_CHANNEL_MAJOR), already projected (emip_linear_cn_fwd)
 *   v     [B,2,N]  (the flow viewed as [B,2,H*W])      out [B,2,N]
 *   lse   NULL or [B,N] (needed by the backward) */
size_t emip_flow_attn_workspace(int B, int N, int C);
int emip_flow_attn_fwd(const float* q, const float* k, const float* v, float* out, float* lse, void* workspace,
                       size_t ws_bytes, int B, int N, int C, int flags, void* stream);
/* a2 as one call on the FeatureTransformer's token rows (transformer.py:519-532; inference, forward only): query = q_proj(x),
 * key = k_proj(query) -- the reference projects the key from the PROJECTED query (:523-524) --, out = softmax(query key^T /
 * sqrt(C)) v.  x_split = the bf16 rows [B][N][hi 128 | lo 128] (emip_feature_transformer_fwd_ex out_split); both projection
 * GEMMs read their row operand pre-split and write query / key as the bf16 hi | lo operands of the attention kernel
 * (bias added in the epilogue), so no fp32 query / key and no split pass exist.  wq, wk [C,C] and bq, bk [C] fp32 (x_split
 * 128-byte, bq / bk 16-byte aligned); v, out [B,2,N]; workspace >= emip_flow_attn_tokens_workspace(B, N, C) bytes, 1024-byte aligned.
 * flags: 0, EMIP_FLAG_BF16, EMIP_FLAG_SCHED_*.  Bit-identical to emip_linear_tm_bias_fwd x 2 + emip_flow_attn_fwd. */
size_t emip_flow_attn_tokens_workspace(int B, int N, int C);
int emip_flow_attn_tokens_fwd(const void* x_split, const float* wq, const float* bq, const float* wk, const float* bk, const float* v,
                              float* out, void* workspace, size_t ws_bytes, int B, int N, int C, int flags, void* stream);
/* Backward w.r.t. q and k only: the value is flow.detach() in the model (gmflow.py:137). */
int emip_flow_attn_bwd(const float* q, const float* k, const float* v, const float* out, const float* lse,
                       const float* dout, float* dq, float* dk, void* workspace, size_t ws_bytes, int B, int N,
                       int C, int flags, void* stream);

/* The two projections of FeatureFlowAttention (transformer.py:523-524) on the feature map as it lies in memory:
 * y[b][m][n] = sum_k w[m][k] x[b][k][n] + bias[m], x [B,K,N], y [B,M,N] (channel-major), w [M,K] as nn.Linear stores it.
 * Forward and backward are split-bf16 tensor-core GEMMs; dw / db may be NULL (frozen weights).  M, K <= 1024, N % 4 == 0. */
size_t emip_linear_cn_workspace(int B, int M, int K, int N);
int emip_linear_cn_fwd(const float* x, const float* w, const float* bias, float* y, void* workspace, size_t ws_bytes, int B, int M,
                       int K, int N, void* stream);
int emip_linear_cn_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, float* db, void* workspace,
                       size_t ws_bytes, int B, int M, int K, int N, void* stream);

/* ---- f2b: token-major layers of the FeatureTransformer blocks ------------------------------------ */
/* Replaces the nn.Linear / nn.GELU / nn.LayerNorm calls of TransformerLayer.forward,
 * model/EMIP_short/motion/gmflow/transformer.py:163-176 (q_proj, k_proj, v_proj :163-165, merge :171, norm1 :172,
 * mlp = Linear(256,1024) - GELU - Linear(1024,128) :175 (defined :140-146), norm2 :176, residual :180); all six linear
 * layers are bias-free (:126-131, :143-147).  Rows are tokens: x [L,K], y [L,M], w [M,K] as nn.Linear stores it.
 *   y[l][m] = sum_k act(x[l][k]) w[m][k]          act = exact GELU with EMIP_LINEAR_GELU_IN, else identity
 *   EMIP_LINEAR_W_TRANS: x [L,M], y [L,K] = act(x) w  (the input gradient of the layer)
 * Split-bf16 tensor-core GEMM (three UMMAs, fp32 accumulation).  M % 4 == 0, K % 4 == 0, K >= 16; workspace 1024-byte aligned. */
#define EMIP_LINEAR_GELU_IN 1
#define EMIP_LINEAR_W_TRANS 2
size_t emip_linear_tm_workspace(int L, int M, int K);
int emip_linear_tm_fwd(const float* x, const float* w, float* y, void* workspace, size_t ws_bytes, int L, int M, int K, int flags,
                       void* stream);
/* n layers of one shape on the same rows (q / k / v of a self-attention layer, k / v of a cross-attention layer,
 * transformer.py:163-165): y[i] = x w[i]^T with the rows split once.  w, y: host arrays of n device pointers.  K % 64 == 0. */
size_t emip_linear_tm_multi_workspace(int L, int M, int K);
int emip_linear_tm_multi_fwd(const float* x, const float* const* w, float* const* y, int n, void* workspace, size_t ws_bytes, int L,
                             int M, int K, void* stream);
/* EMIP_LINEAR_GELU_BWD_IN: the rows are x * GELU'(aux), aux laid out like x (the pre-activation the forward kept) -- with
 * EMIP_LINEAR_W_TRANS this is the input gradient of mlp[0] with the derivative of mlp[1] applied on load. */
#define EMIP_LINEAR_GELU_BWD_IN 4
int emip_linear_tm_fwd_ex(const float* x, const float* aux, const float* w, float* y, void* workspace, size_t ws_bytes, int L, int M,
                          int K, int flags, void* stream);
/* The MLP of a cross-attention block in one call (transformer.py:140-146, :175): y [L,M] = GELU(x [L,K1] w1^T) w2^T with
 * w1 [Hd,K1], w2 [M,Hd].  The first GEMM's epilogue applies the exact GELU and writes the hidden rows as the bf16 hi | lo
 * operand of the second GEMM: no fp32 hidden tensor, no separate split pass (inference / no-grad path; with autograd the
 * host mirror issues two emip_linear_tm_fwd calls and keeps the pre-activation).  Hd % 64 == 0. */
size_t emip_mlp_tm_workspace(int L, int K1, int Hd, int M);
int emip_mlp_tm_fwd(const float* x, const float* w1, const float* w2, float* y, void* workspace, size_t ws_bytes, int L, int K1,
                    int Hd, int M, void* stream);
/* No-grad fast paths with the LayerNorm (+ residual) of transformer.py:172 / :176 / :180 inside the GEMM epilogue (M = 128: one
 * output row per thread, the statistics never leave the registers; the pre-norm tensor never reaches HBM):
 *   emip_linear_ln_tm_fwd: y = res + LN(act(x) w^T) * gamma + beta                    (merge + norm1 [+ source])
 *   emip_mlp_ln_tm_fwd:    y = res + LN(GELU(x w1^T) w2^T) * gamma + beta; gamma NULL = emip_mlp_tm_fwd   (mlp + norm2 + source)
 * res may be NULL; workspaces as emip_linear_tm_workspace / emip_mlp_tm_workspace. */
int emip_linear_ln_tm_fwd(const float* x, const float* w, const float* gamma, const float* beta, const float* res, float* y,
                          void* workspace, size_t ws_bytes, int L, int M, int K, float eps, int flags, void* stream);
int emip_mlp_ln_tm_fwd(const float* x, const float* w1, const float* w2, const float* gamma, const float* beta, const float* res,
                       float* y, void* workspace, size_t ws_bytes, int L, int K1, int Hd, int M, float eps, void* stream);
/* y = (res ? res : 0) + LayerNorm_C(x) * gamma + beta over the channel axis of [L,C] rows (C = 128), biased variance, eps as
 * nn.LayerNorm; the backward returns dx only (frozen affine parameters), statistics are recomputed from x. */
int emip_layernorm_tm_fwd(const float* x, const float* gamma, const float* beta, const float* res, float* y, int L, int C,
                          float eps, void* stream);
int emip_layernorm_tm_bwd(const float* x, const float* gamma, const float* dy, float* dx, int L, int C, float eps, void* stream);

/* ---- a4: prompt fusion (camouflaged feeder / motion collector) ---------------------------------- */
/* Replaces model/EMIP_short/motion/PromptInteract.py:452-464 Injector.forward(image_embeddings, flow)
 * = TransformerBlock_MDTA :436-450 (LayerNorm :333-362, Attention_MDTA :390-432, FeedForward :367-385)
 * and its autograd backward.  dim 128, 2 heads, hidden 340; exact fp32.
 *   x, x1, out, dout, dx, dx1   [B,128,H,W] contiguous
 *   params / dparams            15 device pointers in the order of the reference's state_dict:
 *        transformer.{norm1.body.weight, norm1.body.bias, norm2.body.weight, norm2.body.bias, norm3.body.weight,
 *        norm3.body.bias, attn.temperature, attn.q.weight, attn.q_dwconv.weight, attn.kv.weight,
 *        attn.kv_dwconv.weight, attn.project_out.weight, ffn.project_in.weight, ffn.dwconv.weight,
 *        ffn.project_out.weight}; dparams are overwritten (not accumulated).
 *   saved      >= emip_injector_saved_bytes() : activations written by fwd, read by bwd (caller keeps it alive)
 *   workspace  >= emip_injector_workspace()   : scratch, 256-byte aligned */
size_t emip_injector_saved_bytes(int B, int H, int W);
size_t emip_injector_workspace(int B, int H, int W);
int emip_injector_fwd(const float* x, const float* x1, const float* const* params, float* out, void* saved,
                      size_t saved_bytes, void* workspace, size_t ws_bytes, int B, int H, int W, void* stream);
/* Same with flags: 0 = the five 1x1 convolutions on the tensor cores (3-term split-bf16 operands, fp32 accumulation,
 * rel-L2 ~1e-5), EMIP_FLAG_EXACT_FP32 = exact-fp32 CUDA-core GEMMs.  emip_injector_fwd == flags 0. */
int emip_injector_fwd_ex(const float* x, const float* x1, const float* const* params, float* out, void* saved,
                         size_t saved_bytes, void* workspace, size_t ws_bytes, int B, int H, int W, int flags, void* stream);
int emip_injector_bwd(const float* x, const float* x1, const float* const* params, const void* saved, size_t saved_bytes,
                      const float* dout, float* dx, float* dx1, float* const* dparams, void* workspace, size_t ws_bytes,
                      int B, int H, int W, void* stream);
/* Same with flags (0 = input- and weight-gradient GEMMs on the tensor cores, EMIP_FLAG_EXACT_FP32 = CUDA cores). */
int emip_injector_bwd_ex(const float* x, const float* x1, const float* const* params, const void* saved, size_t saved_bytes,
                         const float* dout, float* dx, float* dx1, float* const* dparams, void* workspace, size_t ws_bytes,
                         int B, int H, int W, int flags, void* stream);

/* ---- a5: EMIP_long memory read (historical-feature prompt) ------------------------------------ */
/* Replaces model/EMIP_long/LTM.py:49-68 Memory.forward(m_in, m_out, q_in, q_out):
 * p = softmax over the memory axis of m_in^T q_in / sqrt(De); mem = m_out p.  Exact fp32, flash-style (p, the
 * reference's unused `viz` output, is never materialised).
 *   m_in, m_out [B,128,M] (M = T*H*W memory slots, the reference's [B,128,T,H,W] viewed flat)   q_in [B,128,Q]
 *   mem  element (b,o,q) at mem[b*mem_stride_b + o*Q + q]: pass the first half of the [B,256,H,W] output of the
 *        reference's torch.cat([mem, q_out]) (LTM.py:66) with mem_stride_b = 256*Q, or a plain [B,128,Q] tensor
 *   lse  [B,Q] log-sum-exp over the memory axis (NULL in inference; required by the backward)
 * Backward: dmem uses the same addressing as mem; dm_in, dm_out [B,128,M] and dq_in [B,128,Q] are overwritten. */
size_t emip_memory_read_workspace(int B, int De, int Do, int M, int Q);
int emip_memory_read_fwd(const float* m_in, const float* m_out, const float* q_in, float* mem, long long mem_stride_b,
                         float* lse, void* workspace, size_t ws_bytes, int B, int De, int Do, int M, int Q, void* stream);
int emip_memory_read_bwd(const float* m_in, const float* m_out, const float* q_in, const float* mem, long long mem_stride_b,
                         const float* lse, const float* dmem, long long dmem_stride_b, float* dm_in, float* dm_out,
                         float* dq_in, void* workspace, size_t ws_bytes, int B, int De, int Do, int M, int Q, void* stream);

/* a5 forward on the tensor cores (csrc/memory_read_tc.cu): same arguments and results as emip_memory_read_fwd
 * (3-term split-bf16 operands, fp32 accumulation: rel-L2 ~1e-5 instead of ~1e-7); M >= 16; workspace of
 * emip_memory_read_tc_workspace() bytes, 1024-byte aligned.  lse (optional) is what emip_memory_read_bwd takes. */
size_t emip_memory_read_tc_workspace(int B, int De, int Do, int M, int Q);
int emip_memory_read_fwd_tc(const float* m_in, const float* m_out, const float* q_in, float* mem, long long mem_stride_b,
                            float* lse, void* workspace, size_t ws_bytes, int B, int De, int Do, int M, int Q, void* stream);

/* a5 backward on the tensor cores (csrc/attn_bwd_tc.cu: one kernel, three launches -- dq_in with rows = queries,
 * dm_in and dm_out with rows = memory slots and the score tile recomputed transposed; 3-term split-bf16 operands, fp32
 * accumulation).  Same arguments and results as emip_memory_read_bwd; workspace of
 * emip_memory_read_bwd_tc_workspace() bytes, 1024-byte aligned. */
size_t emip_memory_read_bwd_tc_workspace(int B, int De, int Do, int M, int Q);
int emip_memory_read_bwd_tc(const float* m_in, const float* m_out, const float* q_in, const float* mem, long long mem_stride_b,
                            const float* lse, const float* dmem, long long dmem_stride_b, float* dm_in, float* dm_out,
                            float* dq_in, void* workspace, size_t ws_bytes, int B, int De, int Do, int M, int Q, void* stream);

/* ---- f4: convex x8 upsampling of the coarse flow -------------------------------- */
/* Replaces model/EMIP_short/motion/gmflow/gmflow.py:64-77, the part of GMFlow.upsample_flow after
 * `mask = self.upsampler(concat)`: softmax over the 9 taps, 3x3 unfold of k*flow, weighted sum, pixel shuffle.
 *   flow [B,2,h,w]   mask [B,9*k*k,h,w] (channel = tap*k*k + ky*k + kx)   out, dout [B,2,k*h,k*w]   k must be 8
 * Backward: dflow [B,2,h,w] and dmask [B,9*k*k,h,w] are overwritten; workspace >= emip_convex_upsample_workspace(). */
size_t emip_convex_upsample_workspace(int B, int h, int w);
int emip_convex_upsample_fwd(const float* flow, const float* mask, float* out, int B, int h, int w, int k, void* stream);
int emip_convex_upsample_bwd(const float* flow, const float* mask, const float* dout, float* dflow, float* dmask,
                             void* workspace, size_t ws_bytes, int B, int h, int w, int k, void* stream);

/* ---- f3: backward-flow occlusion mask of the photometric loss ----------------------- */
/* Replaces loss/warp_utils.py:106-112 get_occu_mask_backward(flow21, th) (get_corresponding_map :26-80): forward
 * splat of every pixel's bilinear weights along flow21, clamp to [0,1], compare with th.  Not differentiable.
 *   flow21 addressed like flow in emip_flow_warp_fwd (channel-slice strides)   mask [B,1,H,W] (1 = occluded)
 *   workspace >= emip_occu_mask_workspace() holds the splat accumulator. */
size_t emip_occu_mask_workspace(int B, int H, int W);
int emip_occu_mask_backward(const float* flow21, float* mask, void* workspace, size_t ws_bytes, int B, int H, int W,
                            long long flow_stride_b, long long flow_stride_c, float th, void* stream);

/* ---- f1: first layer of conv_corr on the never-materialised cost volume --------------- */
/* Replaces model/EMIP_short/model.py:59 (nn.Conv2d(H*W, O, 3, 1, 1), first layer of conv_corr) applied at model.py:96
 * to corr = matching.py:16-20: out[b,o,y,x] = bias[o] + sum_{j,dy,dx} w[o,j,dy,dx] corr[b,j,y+dy-1,x+dx-1] with
 * corr[b,j,y,x] = sum_c f0[b,c,(y,x)] f1[b,c,j] / sqrt(C), computed as two per-sample GEMMs on the feature maps
 * (csrc/conv_corr.cu) so that corr is never formed.
 *   f0, f1 [B,C,H,W] (the matching features)   w [O, H*W, 3, 3]   bias [O] or NULL   out [B,O,H,W]
 * The weight is permuted and split into bf16 hi|lo once per weight version by emip_conv_corr_prepare_weight into a
 * caller-owned buffer of emip_conv_corr_weight_bytes(O, H*W) bytes (1024-byte aligned).
 * emip_conv_corr_supported: C == 128 and some R with R*W % 16 == 0, R*W <= 256 (44x44: R = 4). */
int emip_conv_corr_supported(int C, int H, int W);
size_t emip_conv_corr_weight_bytes(int O, int N);
int emip_conv_corr_prepare_weight(const float* w, void* w_prep, int O, int N, void* stream);
size_t emip_conv_corr_workspace(int B, int C, int H, int W, int O);
int emip_conv_corr_fwd(const float* f0, const float* f1, const void* w_prep, const float* bias, float* out,
                       void* workspace, size_t ws_bytes, int B, int C, int H, int W, int O, void* stream);

/* As emip_conv_corr_fwd with (i) `layout` of f0 / f1 (EMIP_LAYOUT_*) and (ii) the eval-mode BatchNorm2d + ReLU that follow
 * conv_corr[0] (model.py:60-61) folded into the epilogue: ep_scale != NULL => out = act(conv * ep_scale[o] + ep_shift[o])
 * with ep_scale = gamma / sqrt(running_var + eps), ep_shift = (bias - running_mean) * ep_scale + beta (bias is then
 * ignored); relu != 0 => act = ReLU. */
int emip_conv_corr_fwd_ex(const float* f0, const float* f1, const void* w_prep, const float* bias, const float* ep_scale,
                          const float* ep_shift, int relu, int layout, float* out, void* workspace, size_t ws_bytes, int B, int C,
                          int H, int W, int O, void* stream);

/* As emip_conv_corr_fwd_ex; tok_out != NULL (out may then be NULL): the result goes out as the token-major bf16 hi | lo input
 * operand of a following emip_conv3x3_fwd_tokens -- [B][H*W][2*Cp], Cp = O rounded up to 128, hi at [0,Cp), lo at [Cp,2Cp), padding
 * channels zero; 1024-byte aligned -- so conv_corr[0] + BatchNorm + ReLU hand conv_corr[3] its operand with no fp32 round trip. */
int emip_conv_corr_fwd_tokens(const float* f0, const float* f1, const void* w_prep, const float* bias, const float* ep_scale,
                              const float* ep_shift, int relu, int layout, float* out, void* tok_out, void* workspace,
                              size_t ws_bytes, int B, int C, int H, int W, int O, void* stream);

/* Backward of emip_conv_corr_fwd on the tensor cores: df0, df1 [B,C,H,W], dweight [O,H*W,3,3], dbias [O] (NULL = not
 * wanted) from dout [B,O,H,W]; weight = the fp32 parameter, w_prep = its prepared copy.  Five split-bf16 GEMMs
 * (csrc/gemm_tc.cu); workspace of emip_conv_corr_bwd_workspace() bytes, 1024-byte aligned. */
size_t emip_conv_corr_bwd_workspace(int B, int C, int H, int W, int O);
int emip_conv_corr_bwd(const float* f0, const float* f1, const float* weight, const void* w_prep, const float* dout, float* df0,
                       float* df1, float* dweight, float* dbias, void* workspace, size_t ws_bytes, int B, int C, int H, int W,
                       int O, void* stream);

/* ---- f3, second half: photometric term of the unsupervised flow loss, fused -------- */
/* Replaces loss/loss_flow.py:35-49 unFlowLoss.loss_photomatric(im1_scaled, im1_recons, occu_mask1) with
 * loss/loss_blocks.py:46-65 SSIM (3x3, no padding), w_ternary = 0:
 *   loss = (w_l1 mean(|im - rec| m) + w_ssim mean(clamp((1 - SSIM(rec m, im m)) / 2, 0, 1))) / mean(m)
 *   im, rec, drec [B,C,H,W]   mask [B,1,H,W]   loss: device scalar (may be NULL)   sums: device float[4]
 *   (sum |im - rec| m, sum dist, sum m, loss) written by fwd and read by bwd   gloss: device scalar dL/dloss
 * Only rec carries a gradient (loss_flow.py:90-91).  H, W >= 3. */
size_t emip_photometric_workspace(int B, int H, int W);
int emip_photometric_fwd(const float* im, const float* rec, const float* mask, float* loss, float* sums, void* workspace,
                         size_t ws_bytes, int B, int C, int H, int W, float w_l1, float w_ssim, void* stream);
int emip_photometric_bwd(const float* im, const float* rec, const float* mask, const float* sums, const float* gloss,
                         float* drec, int B, int C, int H, int W, float w_l1, float w_ssim, void* stream);

/* ---- f2: attention core of the GMFlow FeatureTransformer ----------------------------- */
/* Replaces model/EMIP_short/motion/gmflow/transformer.py:8-16 single_head_full_attention and the per-window attention
 * of :46-105 single_head_split_window_attention: out = softmax(q k^T / sqrt(C)) v for nb independent problems of n
 * tokens (the shifted-window mask is handled by the host side: it only separates rectangular token blocks, each of
 * which is a plain attention problem).  q, k, v, out token-major [nb][n][C], C = 128. */
size_t emip_attention_tc_workspace(int nb, int n, int C);
int emip_attention_fwd_tc(const float* q, const float* k, const float* v, float* out, void* workspace, size_t ws_bytes,
                          int nb, int n, int C, void* stream);

/* Backward of emip_attention_fwd_tc on the tensor cores: dq, dk, dv [nb][n][C] from q, k, v and dout (self-contained: the
 * fused forward is re-run inside for the output and the row log-sum-exp, nothing has to be saved). */
size_t emip_attention_bwd_tc_workspace(int nb, int n, int C);
int emip_attention_bwd_tc(const float* q, const float* k, const float* v, const float* dout, float* dq, float* dk, float* dv,
                          void* workspace, size_t ws_bytes, int nb, int n, int C, void* stream);

/* One call per FeatureTransformer attention layer (csrc/window_attn.cu): q, k, v, out [B][h*w][C] as the reference's
 * single_head_split_window_attention(q, k, v, num_splits, with_shift, h, w, attn_mask) takes and returns them.  The
 * window partition (and, when shifted, the block structure its 0 / -100 mask implies) is address arithmetic inside the
 * operand-split pass and the attention epilogue: no gather / scatter / roll copies, no mask tensor.  C = 128, at most
 * 16 blocks with the same token count (num_splits <= 3). */
size_t emip_window_attention_tc_workspace(int B, int h, int w, int C, int num_splits, int with_shift);
int emip_window_attention_fwd_tc(const float* q, const float* k, const float* v, float* out, float* lse, void* workspace,
                                 size_t ws_bytes, int B, int h, int w, int C, int num_splits, int with_shift, void* stream);

/* Backward of the call above: dq, dk, dv [B][h*w][C] (every element is written) from q, k, v, the forward's out and lse
 * ([B][h*w] row log-sum-exp in pixel order; pass a buffer as `lse` to the forward, NULL in inference) and dout.  Window
 * gather in the operand-split pass, scatter in the gradient kernels' epilogues. */
size_t emip_window_attention_bwd_tc_workspace(int B, int h, int w, int C, int num_splits, int with_shift);
int emip_window_attention_bwd_tc(const float* q, const float* k, const float* v, const float* out, const float* lse,
                                 const float* dout, float* dq, float* dk, float* dv, void* workspace, size_t ws_bytes, int B,
                                 int h, int w, int C, int num_splits, int with_shift, void* stream);

/* EMIP_WINATTN_KV_SWAP_HALVES: image i attends to the keys / values of image (i + B/2) % B: the cross-attention layers of
 * the FeatureTransformer, whose `target` is `source` with the two batch halves swapped (transformer.py:462, :473) -- the
 * swapped copy is never made.  Forward only. */
#define EMIP_WINATTN_KV_SWAP_HALVES 1
int emip_window_attention_fwd_tc_ex(const float* q, const float* k, const float* v, float* out, float* lse, void* workspace,
                                    size_t ws_bytes, int B, int h, int w, int C, int num_splits, int with_shift, int flags,
                                    void* stream);

/* The whole FeatureTransformer (transformer.py:433-482: n_blocks x (self-attention layer, cross-attention + FFN layer)) as ONE
 * call on token rows, inference only (csrc/window_attn.cu, second half): x, out [B][h*w][128] with B = 2 x pairs (frame 1 |
 * frame 2 on the batch axis, transformer.py:461); out may alias x; odd blocks use shifted windows (:425).
 *   weights  n_blocks x 16 device pointers, per block in the reference's state_dict order: self_attn.{q_proj,k_proj,v_proj,merge}
 *            .weight [128,128], self_attn.norm1.{weight,bias}, cross_attn_ffn.{q_proj,k_proj,v_proj,merge}.weight,
 *            .norm1.{weight,bias}, .mlp.0.weight [1024,256], .mlp.2.weight [128,1024], .norm2.{weight,bias}
 *   prep     emip_feature_transformer_prepare(weights) in a caller-owned buffer of emip_feature_transformer_weight_bytes()
 *            bytes (bf16 hi | lo matrices; redo when a weight changes); the LayerNorm vectors are read in place
 *   workspace >= emip_feature_transformer_workspace(B, h, w, 128) bytes, 1024-byte aligned.
 * Every intermediate is written by its producer's epilogue in the form its consumer reads (window-ordered bf16 hi | lo
 * q / k / v, pre-split merge / MLP operands, LayerNorm + residual inside the GEMM epilogues); the batch-swapped `concat1`
 * (transformer.py:462, :473) and torch.cat([source, message]) (:175) are never materialised. */
size_t emip_feature_transformer_weight_bytes(int n_blocks);
int emip_feature_transformer_prepare(const float* const* weights, int n_blocks, void* prep, void* stream);
size_t emip_feature_transformer_workspace(int B, int h, int w, int C);
int emip_feature_transformer_fwd(const float* x, float* out, const float* const* weights, const void* prep, int n_blocks,
                                 void* workspace, size_t ws_bytes, int B, int h, int w, int C, int num_splits, float eps,
                                 void* stream);
/* Same, plus out_split (may be NULL): the output rows once more as bf16 [B][h*w][hi 128 | lo 128] (128-byte aligned), written by
 * the last block's LayerNorm epilogue -- the operand emip_global_matching_fwd (EMIP_FLAG_PRESPLIT) and
 * emip_flow_attn_tokens_fwd read, so that no split pass over `out` runs (gmflow.py:117-137 after the transformer). */
int emip_feature_transformer_fwd_ex(const float* x, float* out, void* out_split, const float* const* weights, const void* prep,
                                    int n_blocks, void* workspace, size_t ws_bytes, int B, int h, int w, int C, int num_splits,
                                    float eps, void* stream);

/* ---- the chained path between the backbones and the decoder (CoUpdater.forward, model.py:92-97) -------------------- */
/* 3x3 convolution, stride 1, zero padding 1, on the tensor cores (csrc/conv_tm.cu): replaces GMFlow.upsampler[0] + ReLU
 * (gmflow.py:43-44 on cat(flow, feature), :62-64) and conv_corr[3] (model.py:62).
 *   input  = channel concatenation of x0 (C0 channels) and, optionally, x1 (C1 channels); each source is channel-major
 *            [B,C,H,W] or token-major [B,H*W,C] (layout0 / layout1 = EMIP_LAYOUT_*)
 *   w_prep = emip_conv3x3_prepare_weight(w [O, C0+C1, 3, 3]) in a caller-owned buffer of emip_conv3x3_weight_bytes() bytes
 *   out [B,O,H,W] = act(scale[o] * conv + shift[o]); scale NULL = 1, shift NULL = 0 (shift = the bias), relu != 0 = ReLU
 *   workspace >= emip_conv3x3_workspace(B, C0+C1, H, W) bytes, 1024-byte aligned (the bf16 hi | lo token rows). */
int emip_conv3x3_supported(int Cin, int H, int W);
size_t emip_conv3x3_weight_bytes(int O, int Cin);
int emip_conv3x3_prepare_weight(const float* w, void* w_prep, int O, int Cin, void* stream);
size_t emip_conv3x3_workspace(int B, int Cin, int H, int W);
int emip_conv3x3_fwd(const float* x0, int C0, int layout0, const float* x1, int C1, int layout1, const void* w_prep,
                     const float* scale, const float* shift, int relu, float* out, void* workspace, size_t ws_bytes, int B,
                     int H, int W, int O, void* stream);

/* emip_conv3x3_fwd on an input that already is the token-major bf16 hi | lo operand [B][H*W][2*Cp_in] (Cp_in = Cin rounded up to
 * 64; what emip_conv_corr_fwd_tokens writes): no split pass, no workspace. */
int emip_conv3x3_fwd_tokens(const void* tok, int Cin, int Cp_in, const void* w_prep, const float* scale, const float* shift, int relu,
                            float* out, int B, int H, int W, int O, void* stream);

/* out[b][n][c] = x[b][c][n] + pos[c][n] (pos NULL: plain transpose): gmflow.py:114 feature_add_position (utils.py:66-86; pos
 * = the window-tiled sine embedding) + transformer.py:439-440 (flatten / permute to token rows) in one pass. */
int emip_tokens_from_cn(const float* x, const float* pos, float* out, int B, int C, int N, void* stream);

/* y [L][M] = x [L][K] w^T + bias [M]: nn.Linear with bias on token rows -- FeatureFlowAttention.q_proj / k_proj
 * (transformer.py:523-524) applied to the FeatureTransformer's token-major output.  Forward only; workspace as
 * emip_linear_tm_workspace(L, M, K). */
int emip_linear_tm_bias_fwd(const float* x, const float* w, const float* bias, float* y, void* workspace, size_t ws_bytes, int L,
                            int M, int K, void* stream);

/* Diagnostics: device buffer of (CTAs x 16) cycle counters filled by the next fused-attention launches (NULL = off). */
void emip_attn_tc_set_profile_buffer(unsigned long long* dev_buf);

#ifdef __cplusplus
}
#endif
#endif /* EMIP_B200_H */
