// Tensor-core (tcgen05 / TMEM / TMA) forward of the pairwise-softmax family.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

struct MatchTcArgs {
  const void* x_split;     // bf16 [nbx][nq][256]  (hi 128 | lo 128), rows of the score matrix
  const void* y_split;     // bf16 [nby][nk][256], columns of the score matrix
  int nbx, nby;
  // value 2-vector per column.  grid_w > 0: analytic pixel grid (x = col % grid_w, y = col / grid_w), `v` unused.
  const float* v;          // v[p*v_stride_b + ch*nk + col]
  long long v_stride_b;
  int grid_w;
  int sub_grid;            // subtract the row's own grid position from the expectation (matching.py:39)
  float* out;              // [nb][2][nq]
  float* lse;              // optional [nb][nq]
  int nb, nq, nk;
  int y_shift, y_mod;      // problem p: rows X[p], columns Y[(p + y_shift) % y_mod]
  float* s_out;            // optional scaled scores for problems [s_first, s_first+s_count): [s_count][nq][nk]
  int s_first, s_count;
  float sqrt_c;
  int terms;               // 3 = bf16 hi/lo split (hi.hi + lo.hi + hi.lo, fp32-accurate); 1 = single-pass bf16
  int schedule;            // 0 = choose by batch size, 1 = stream-K spans, 2 = grid-strided whole items (EMIP_FLAG_SCHED_*)
  void* sk_ws;             // stream-K bookkeeping, match_tc_streamk_bytes(nb, nq, nk) bytes, 16-byte aligned
  size_t sk_bytes;
};

bool match_tc_supported(int nq, int nk, int c);
size_t match_tc_split_bytes(int nb, int n, int c);
size_t match_tc_streamk_bytes(int nb, int nq, int nk);
// fp32 -> (bf16 hi | bf16 lo) token-major operands; writes batches [dst_batch0, dst_batch0+nb) of dst.
// src2 != NULL: a second source of nb batches is written right behind the first (one launch for f0 and f1).
int match_tc_split(const float* src, const float* src2, void* dst, int nb, int n, int c, int layout, int dst_batch0,
                   cudaStream_t st);
int match_tc_fwd(const MatchTcArgs& a, cudaStream_t st);
