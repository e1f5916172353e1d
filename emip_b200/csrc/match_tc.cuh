// Tensor-core (tcgen05 / TMEM / TMA) forward of the pairwise-softmax family.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

struct MatchTcArgs {
  const void* x_split;     // bf16 [nbx][nq][256]  (hi 128 | lo 128), rows of the score matrix
  const void* y_split;     // bf16 [nby][nk][256], columns of the score matrix
  int nbx, nby;
  const float* v;          // value 2-vectors per column: v[p*v_stride_b + ch*nk + col]
  long long v_stride_b;    // 0 = shared by all problems
  const float* sub;        // optional [2][nq], subtracted from the expectation
  float* out;              // [nb][2][nq]
  float* lse;              // optional [nb][nq]
  int nb, nq, nk;
  int y_shift, y_mod;      // problem p: rows X[p], columns Y[(p + y_shift) % y_mod]
  float* s_out;            // optional scaled scores for problems [s_first, s_first+s_count): [s_count][nq][nk]
  int s_first, s_count;
  float sqrt_c;
};

bool match_tc_supported(int nq, int nk, int c);
size_t match_tc_split_bytes(int nb, int n, int c);
// fp32 -> (bf16 hi | bf16 lo) token-major operands; writes batches [dst_batch0, dst_batch0+nb) of dst
int match_tc_split(const float* src, void* dst, int nb, int n, int c, int layout, int dst_batch0, cudaStream_t st);
int match_tc_fwd(const MatchTcArgs& a, cudaStream_t st);
