// Argument blocks shared by the pairwise-softmax kernels (pair_simt.cu, match_tc.cu)
// and their C-ABI front ends (matching.cu, flow_attn.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#define EMIP_LAYOUT_NC 0   // token-major  [N][C]
#define EMIP_LAYOUT_CN 1   // channel-major [C][N]  (NCHW feature maps)

// Problem p in [0, nb): rows from X[p], columns from Y[(p + y_shift) % nb].
struct PairFwdArgs {
  const float* x;          // [nb][nq*128] in x_layout
  const float* y;          // [nb][nk*128] in y_layout
  const float* v;          // value 2-vectors per column: v[p*v_stride_b + ch*nk + col]
  long long v_stride_b;    // 0 = shared by all problems (the coordinate grid)
  const float* sub;        // optional [2][nq]: subtracted from the expectation (init grid), or NULL
  float* out;              // [nb][2][nq]
  float* lse;              // optional [nb][nq] log-sum-exp of the scaled scores
  float* s_out;            // optional [nb][nq][nk] scaled scores
  int nb, nq, nk;
  int y_shift;
  int x_layout, y_layout;
  float sqrt_c;            // scores are divided by this
};

// dX for problem p: rows from X[(p + x_shift) % nb], columns from Y[(p + y_shift) % nb];
// per-row / per-column terms are indexed by p.  See pair_simt.cu for the formula.
struct PairBwdArgs {
  const float* x;          // [nb][nr*128]
  const float* y;          // [nb][nc*128]
  // term 1 (softmax over columns), enabled when l1 != NULL
  const float* l1;         // [nb][nr]
  const float* u;          // [nb][2][nr]
  const float* u0;         // [nb][nr]
  const float* t;          // t[p*t_stride_b + ch*nc + col]
  long long t_stride_b;
  // term 2 (softmax over rows), enabled when l2 != NULL
  const float* l2;         // [nb][nc]
  const float* w;          // [nb][2][nc]
  const float* w0;         // [nb][nc]
  const float* t2;         // t2[p*t2_stride_b + ch*nr + row]
  long long t2_stride_b;
  // direct gradient on the scaled scores, enabled when e != NULL
  const float* e;          // e[p*e_stride_b + row*e_stride_r + col*e_stride_c]
  long long e_stride_b, e_stride_r, e_stride_c;
  float* dx;               // [nb][nr*128] in dx_layout
  int nb, nr, nc;
  int x_shift, y_shift;
  int x_layout, y_layout, dx_layout;
  int accumulate;          // add into dx instead of overwriting
  float sqrt_c;
};

size_t pair_fwd_simt_smem();
size_t pair_bwd_simt_smem();
int pair_fwd_simt(const PairFwdArgs& a, cudaStream_t st);
int pair_bwd_simt(const PairBwdArgs& a, cudaStream_t st);
int launch_rowdot2(const float* u, const float* o, const float* g, float* out, int nb, int n, cudaStream_t st);
int launch_coords_grid(float* g, int h, int w, cudaStream_t st);
