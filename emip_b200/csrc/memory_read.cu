// a5: EMIP_long memory read (the "historical-feature prompt") -- forward and backward.
//
// Replaces reference model/EMIP_long/LTM.py:49-68 Memory.forward(m_in, m_out, q_in, q_out):
//   p[b,m,q] = softmax over the memory axis m of  sum_c m_in[b,c,m] q_in[b,c,q] / sqrt(De)
//   mem[b,o,q] = sum_m m_out[b,o,m] p[b,m,q] ;  out = cat(mem, q_out)
// i.e. attention with queries = the current frame's key map, keys/values = the memory (M = T*H*W <= 5*1936 slots),
// De = Do = 128, everything channel-major ([C][tokens], the NCHW layout the model hands over).  Exact fp32 on the
// CUDA cores: flash-style (the M x Q probability volume, 75 MB at T = 5, is never stored; the reference returns it
// as `viz` but no caller reads it, model_long.py / LTM.py:129).  The model runs this with B = 1, so the memory axis
// is split across CTAs (partial softmax states merged by a small second kernel) to fill the chip.
#include "common.cuh"
#include "../../include/emip_b200.h"
#include "simt_tiles.cuh"
#include <math.h>

namespace {
using namespace simt;

struct MemArgs {
  const float* q_in;    // [B][128][Q]
  const float* m_in;    // [B][128][M]
  const float* m_out;   // [B][128][M]
  int B, Q, M, nsplit, chunk;   // chunk = memory slots per split (multiple of BN)
  float inv_sqrt_d;
  // forward
  float* part_o;        // [nsplit][B][128][Q] un-normalised partial outputs
  float* part_m;        // [nsplit][B][Q] running max
  float* part_l;        // [nsplit][B][Q] running sum
  // backward
  const float* lse;     // [B][Q]
  const float* dmem;    // dmem[b*dmem_stride_b + o*Q + q]
  long long dmem_stride_b;
  const float* dvec;    // [B][Q]  D_q = sum_o dmem[o][q] mem[o][q]
  float* part_dq;       // [nsplit][B][128][Q]
  float* dm_in;         // [B][128][M]
  float* dm_out;        // [B][128][M]
};

// ---- forward: one CTA = (sample, 64 queries, one split of the memory axis)
__global__ void __launch_bounds__(NT) mem_fwd_kernel(MemArgs p) {
  extern __shared__ __align__(16) float smem[];
  float* Xs = smem;                  // queries   [64 q][LDX]
  float* Ys = Xs + BM * LDX;         // keys      [64 m][LDX]
  float* Vs = Ys + BN * LDX;         // values    [64 m][LDX]
  float* Ws = Vs + BN * LDX;         // P tile    [64 q][LDW]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int nqb = (p.Q + BM - 1) / BM;
  int t = blockIdx.x;
  const int qb = t % nqb; t /= nqb;
  const int sp = t % p.nsplit;
  const int b = t / p.nsplit;
  const int row0 = qb * BM;
  const int m_beg = sp * p.chunk, m_end = min(p.M, m_beg + p.chunk);
  const float* Qp = p.q_in + (size_t)b * KC * p.Q;
  const float* Kp = p.m_in + (size_t)b * KC * p.M;
  const float* Vp = p.m_out + (size_t)b * KC * p.M;
  load_tile(Xs, Qp, 1, row0, p.Q, p.Q);
  float m_run[4], l_run[4], oacc[4][8];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    m_run[a] = -INFINITY;
    l_run[a] = 0.f;
#pragma unroll
    for (int d = 0; d < 8; ++d) oacc[a][d] = 0.f;
  }
  for (int col0 = m_beg; col0 < m_end; col0 += BN) {
    __syncthreads();
    load_tile(Ys, Kp, 1, col0, m_end, p.M);
    load_tile(Vs, Vp, 1, col0, m_end, p.M);
    __syncthreads();
    float acc[4][4];
    s_tile(Xs, Ys, tx, ty, acc);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      float tmax = -INFINITY;
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        const int col = col0 + tx + 16 * bb;
        const float s = (col < m_end) ? acc[a][bb] * p.inv_sqrt_d : -INFINITY;     // LTM.py:59
        acc[a][bb] = s;
        tmax = fmaxf(tmax, s);
      }
      tmax = group16_max(tmax);
      const float m_new = fmaxf(m_run[a], tmax);
      const float corr = expf(m_run[a] - m_new);
      float ls = 0.f;
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        const float pr = expf(acc[a][bb] - m_new);
        ls += pr;
        Ws[(ty + 16 * a) * LDW + tx + 16 * bb] = pr;
      }
      l_run[a] = l_run[a] * corr + ls;
      m_run[a] = m_new;
#pragma unroll
      for (int d = 0; d < 8; ++d) oacc[a][d] *= corr;
    }
    __syncthreads();
    wy_accumulate(Ws, Vs, tx, ty, oacc);            // O[q][o] += sum_m P[q][m] V[m][o]
  }
  // partial state of this split
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const float l = group16_sum(l_run[a]);
    const int row = row0 + ty + 16 * a;
    if (tx == 0 && row < p.Q) {
      p.part_m[((size_t)sp * p.B + b) * p.Q + row] = m_run[a];
      p.part_l[((size_t)sp * p.B + b) * p.Q + row] = l;
    }
  }
  __syncthreads();
  store_tile_cn(Xs, oacc, 1.0f, p.part_o + ((size_t)sp * p.B + b) * KC * p.Q, p.Q, row0, p.Q, tx, ty, false);
}

// merge the splits: out[b][o][q] = sum_s e^{m_s - m} O_s / sum_s e^{m_s - m} l_s ; lse = m + log(sum)
__global__ void mem_merge_kernel(const float* __restrict__ part_o, const float* __restrict__ part_m,
                                 const float* __restrict__ part_l, float* __restrict__ out, long long out_stride_b,
                                 float* __restrict__ lse, int B, int Q, int nsplit) {
  const int b = blockIdx.y;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  float m = -INFINITY;
  for (int s = 0; s < nsplit; ++s) m = fmaxf(m, part_m[((size_t)s * B + b) * Q + q]);
  float l = 0.f;
  for (int s = 0; s < nsplit; ++s) l += expf(part_m[((size_t)s * B + b) * Q + q] - m) * part_l[((size_t)s * B + b) * Q + q];
  if (lse != nullptr) lse[(size_t)b * Q + q] = m + logf(l);
  const float inv = 1.0f / l;
  for (int o = 0; o < KC; ++o) {
    float acc = 0.f;
    for (int s = 0; s < nsplit; ++s)
      acc += expf(part_m[((size_t)s * B + b) * Q + q] - m) * part_o[(((size_t)s * B + b) * KC + o) * Q + q];
    out[(size_t)b * out_stride_b + (size_t)o * Q + q] = acc * inv;
  }
}

// D_q = sum_o dmem[o][q] mem[o][q]
__global__ void mem_dvec_kernel(const float* __restrict__ dmem, long long dmem_stride_b, const float* __restrict__ mem,
                                long long mem_stride_b, float* __restrict__ dvec, int Q) {
  const int b = blockIdx.y;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  float s = 0.f;
  for (int o = 0; o < KC; ++o)
    s = fmaf(__ldg(dmem + (size_t)b * dmem_stride_b + (size_t)o * Q + q), __ldg(mem + (size_t)b * mem_stride_b + (size_t)o * Q + q), s);
  dvec[(size_t)b * Q + q] = s;
}

// ---- backward w.r.t. the queries: rows = queries, loop over a split of the memory
//   P = exp(S - lse) ; dP[q][m] = sum_o dmem[o][q] V[o][m] ; W = P (dP - D_q) ; dq_in[c][q] += sum_m W[q][m] K[c][m] / sqrt(De)
__global__ void __launch_bounds__(NT) mem_bwd_q_kernel(MemArgs p) {
  extern __shared__ __align__(16) float smem[];
  float* Xs = smem;                  // queries [64][LDX]
  float* Gs = Xs + BM * LDX;         // dmem rows [64 q][128 o]
  float* Ys = Gs + BM * LDX;         // keys
  float* Vs = Ys + BN * LDX;         // values
  float* Ws = Vs + BN * LDX;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int nqb = (p.Q + BM - 1) / BM;
  int t = blockIdx.x;
  const int qb = t % nqb; t /= nqb;
  const int sp = t % p.nsplit;
  const int b = t / p.nsplit;
  const int row0 = qb * BM;
  const int m_beg = sp * p.chunk, m_end = min(p.M, m_beg + p.chunk);
  load_tile(Xs, p.q_in + (size_t)b * KC * p.Q, 1, row0, p.Q, p.Q);
  load_tile(Gs, p.dmem + (size_t)b * p.dmem_stride_b, 1, row0, p.Q, p.Q);
  float rl[4], rd[4], dacc[4][8];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int row = row0 + ty + 16 * a;
    rl[a] = row < p.Q ? __ldg(p.lse + (size_t)b * p.Q + row) : 0.f;
    rd[a] = row < p.Q ? __ldg(p.dvec + (size_t)b * p.Q + row) : 0.f;
#pragma unroll
    for (int d = 0; d < 8; ++d) dacc[a][d] = 0.f;
  }
  for (int col0 = m_beg; col0 < m_end; col0 += BN) {
    __syncthreads();
    load_tile(Ys, p.m_in + (size_t)b * KC * p.M, 1, col0, m_end, p.M);
    load_tile(Vs, p.m_out + (size_t)b * KC * p.M, 1, col0, m_end, p.M);
    __syncthreads();
    float acc[4][4], dp[4][4];
    s_tile(Xs, Ys, tx, ty, acc);
    s_tile(Gs, Vs, tx, ty, dp);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int row = row0 + ty + 16 * a;
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        const int col = col0 + tx + 16 * bb;
        float w = expf(acc[a][bb] * p.inv_sqrt_d - rl[a]) * (dp[a][bb] - rd[a]);
        if (row >= p.Q || col >= m_end) w = 0.f;
        Ws[(ty + 16 * a) * LDW + tx + 16 * bb] = w;
      }
    }
    __syncthreads();
    wy_accumulate(Ws, Ys, tx, ty, dacc);
  }
  __syncthreads();
  store_tile_cn(Xs, dacc, p.inv_sqrt_d, p.part_dq + ((size_t)sp * p.B + b) * KC * p.Q, p.Q, row0, p.Q, tx, ty, false);
}

// ---- backward w.r.t. the memory: rows = memory slots, loop over all queries
//   dm_in[c][m] = sum_q W[q][m] q_in[c][q] / sqrt(De) ;  dm_out[o][m] = sum_q P[q][m] dmem[o][q]
__global__ void __launch_bounds__(NT) mem_bwd_m_kernel(MemArgs p) {
  extern __shared__ __align__(16) float smem[];
  float* Ks = smem;                  // keys   [64 m][LDX]
  float* Vs = Ks + BM * LDX;         // values [64 m][LDX]
  float* Xs = Vs + BM * LDX;         // queries [64 q][LDX]
  float* Gs = Xs + BN * LDX;         // dmem    [64 q][LDX]
  float* Ws = Gs + BN * LDX;         // W tile  [64 m][LDW]
  float* Ps = Ws + BM * LDW;         // P tile  [64 m][LDW]
  float* Cs = Ps + BM * LDW;         // [2][BN]: lse, D per query column
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int nmb = (p.M + BM - 1) / BM;
  const int b = blockIdx.x / nmb, mb = blockIdx.x % nmb;
  const int row0 = mb * BM;
  load_tile(Ks, p.m_in + (size_t)b * KC * p.M, 1, row0, p.M, p.M);
  load_tile(Vs, p.m_out + (size_t)b * KC * p.M, 1, row0, p.M, p.M);
  float dk[4][8], dv[4][8];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int d = 0; d < 8; ++d) { dk[a][d] = 0.f; dv[a][d] = 0.f; }
  for (int col0 = 0; col0 < p.Q; col0 += BN) {
    __syncthreads();
    load_tile(Xs, p.q_in + (size_t)b * KC * p.Q, 1, col0, p.Q, p.Q);
    load_tile(Gs, p.dmem + (size_t)b * p.dmem_stride_b, 1, col0, p.Q, p.Q);
    if (tid < 2 * BN) {
      const int k = tid / BN, c = tid % BN, col = col0 + c;
      Cs[tid] = col < p.Q ? __ldg((k == 0 ? p.lse : p.dvec) + (size_t)b * p.Q + col) : 0.f;
    }
    __syncthreads();
    float acc[4][4], dp[4][4];
    s_tile(Ks, Xs, tx, ty, acc);      // S^T[m][q]
    s_tile(Vs, Gs, tx, ty, dp);       // dP^T[m][q]
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int row = row0 + ty + 16 * a;
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        const int c = tx + 16 * bb, col = col0 + c;
        float pr = expf(acc[a][bb] * p.inv_sqrt_d - Cs[c]);
        if (row >= p.M || col >= p.Q) pr = 0.f;
        Ps[(ty + 16 * a) * LDW + c] = pr;
        Ws[(ty + 16 * a) * LDW + c] = pr * (dp[a][bb] - Cs[BN + c]);
      }
    }
    __syncthreads();
    wy_accumulate(Ws, Xs, tx, ty, dk);
    wy_accumulate(Ps, Gs, tx, ty, dv);
  }
  __syncthreads();
  store_tile_cn(Xs, dk, p.inv_sqrt_d, p.dm_in + (size_t)b * KC * p.M, p.M, row0, p.M, tx, ty, false);
  __syncthreads();
  store_tile_cn(Xs, dv, 1.0f, p.dm_out + (size_t)b * KC * p.M, p.M, row0, p.M, tx, ty, false);
}

__global__ void sum_parts_kernel(const float* __restrict__ part, long long stride, float* __restrict__ out, int n_parts,
                                 long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int k = 0; k < n_parts; ++k) s += __ldg(part + (size_t)k * stride + i);
  out[i] = s;
}

int pick_split(int B, int Q, int M, int* chunk) {
  const int nqb = (Q + BM - 1) / BM, ntile = (M + BN - 1) / BN;
  int ns = (emip_num_sms() + B * nqb - 1) / (B * nqb);
  if (ns < 1) ns = 1;
  if (ns > ntile) ns = ntile;
  if (ns > 8) ns = 8;
  const int tiles_per = (ntile + ns - 1) / ns;
  *chunk = tiles_per * BN;
  return (ntile + tiles_per - 1) / tiles_per;
}
size_t al(size_t n_floats) { return emip_align_up(n_floats * sizeof(float), 256); }
constexpr size_t SMEM_FWD = sizeof(float) * (3 * BM * LDX + BM * LDW);
constexpr size_t SMEM_BQ = sizeof(float) * (4 * BM * LDX + BM * LDW);
constexpr size_t SMEM_BM = sizeof(float) * (4 * BM * LDX + 2 * BM * LDW + 2 * BN);

}  // namespace

extern "C" size_t emip_memory_read_workspace(int B, int De, int Do, int M, int Q) {
  if (B < 0 || M <= 0 || Q <= 0) return 0;
  // 8 splits at most: partial outputs / dq, partial (m, l), D vector
  return 8 * al((size_t)B * KC * Q) + 16 * al((size_t)B * Q) + al((size_t)B * Q);
}

extern "C" int emip_memory_read_fwd(const float* m_in, const float* m_out, const float* q_in, float* mem, long long mem_stride_b,
                                    float* lse, void* workspace, size_t ws_bytes, int B, int De, int Do, int M, int Q,
                                    void* stream) {
  if (B == 0) return EMIP_OK;
  EMIP_CHECK_ARG(m_in && m_out && q_in && mem && workspace, "memory_read_fwd: null pointer");
  EMIP_CHECK_ARG(M > 0 && Q > 0, "memory_read_fwd: bad shape M=%d Q=%d", M, Q);
  if (De != KC || Do != KC) {
    emip_set_error("memory_read_fwd: De=%d Do=%d unsupported (kernels are built for the model's 128/128)", De, Do);
    return EMIP_ENOSYS;
  }
  if (ws_bytes < emip_memory_read_workspace(B, De, Do, M, Q)) {
    emip_set_error("memory_read_fwd: workspace too small");
    return EMIP_ENOMEM;
  }
  cudaStream_t st = (cudaStream_t)stream;
  MemArgs p = {};
  p.q_in = q_in; p.m_in = m_in; p.m_out = m_out; p.B = B; p.Q = Q; p.M = M;
  p.nsplit = pick_split(B, Q, M, &p.chunk);
  p.inv_sqrt_d = 1.0f / sqrtf((float)De);
  char* w = static_cast<char*>(workspace);
  p.part_o = reinterpret_cast<float*>(w); w += 8 * al((size_t)B * KC * Q);
  p.part_m = reinterpret_cast<float*>(w); w += 8 * al((size_t)B * Q);
  p.part_l = reinterpret_cast<float*>(w);
  if (int rc__ = emip_func_max_smem((const void*)(mem_fwd_kernel), (int)SMEM_FWD)) return rc__;
  const int nqb = (Q + BM - 1) / BM;
  mem_fwd_kernel<<<B * p.nsplit * nqb, NT, SMEM_FWD, st>>>(p);
  EMIP_CHECK_LAUNCH("mem_fwd");
  mem_merge_kernel<<<dim3((Q + 127) / 128, B), 128, 0, st>>>(p.part_o, p.part_m, p.part_l, mem, mem_stride_b, lse, B, Q,
                                                            p.nsplit);
  EMIP_CHECK_LAUNCH("mem_merge");
  return EMIP_OK;
}

extern "C" int emip_memory_read_bwd(const float* m_in, const float* m_out, const float* q_in, const float* mem,
                                    long long mem_stride_b, const float* lse, const float* dmem, long long dmem_stride_b,
                                    float* dm_in, float* dm_out, float* dq_in, void* workspace, size_t ws_bytes, int B,
                                    int De, int Do, int M, int Q, void* stream) {
  if (B == 0) return EMIP_OK;
  EMIP_CHECK_ARG(m_in && m_out && q_in && mem && lse && dmem && dm_in && dm_out && dq_in && workspace,
                 "memory_read_bwd: null pointer");
  if (De != KC || Do != KC) {
    emip_set_error("memory_read_bwd: De=%d Do=%d unsupported", De, Do);
    return EMIP_ENOSYS;
  }
  if (ws_bytes < emip_memory_read_workspace(B, De, Do, M, Q)) {
    emip_set_error("memory_read_bwd: workspace too small");
    return EMIP_ENOMEM;
  }
  cudaStream_t st = (cudaStream_t)stream;
  MemArgs p = {};
  p.q_in = q_in; p.m_in = m_in; p.m_out = m_out; p.B = B; p.Q = Q; p.M = M;
  p.nsplit = pick_split(B, Q, M, &p.chunk);
  p.inv_sqrt_d = 1.0f / sqrtf((float)De);
  p.lse = lse; p.dmem = dmem; p.dmem_stride_b = dmem_stride_b; p.dm_in = dm_in; p.dm_out = dm_out;
  char* w = static_cast<char*>(workspace);
  p.part_dq = reinterpret_cast<float*>(w); w += 8 * al((size_t)B * KC * Q) + 16 * al((size_t)B * Q);
  float* dvec = reinterpret_cast<float*>(w);
  p.dvec = dvec;
  if (int rc__ = emip_func_max_smem((const void*)(mem_bwd_q_kernel), (int)SMEM_BQ)) return rc__;
  if (int rc__ = emip_func_max_smem((const void*)(mem_bwd_m_kernel), (int)SMEM_BM)) return rc__;
  mem_dvec_kernel<<<dim3((Q + 127) / 128, B), 128, 0, st>>>(dmem, dmem_stride_b, mem, mem_stride_b, dvec, Q);
  EMIP_CHECK_LAUNCH("mem_dvec");
  const int nqb = (Q + BM - 1) / BM, nmb = (M + BM - 1) / BM;
  mem_bwd_q_kernel<<<B * p.nsplit * nqb, NT, SMEM_BQ, st>>>(p);
  EMIP_CHECK_LAUNCH("mem_bwd_q");
  const long long n = (long long)B * KC * Q;
  sum_parts_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p.part_dq, n, dq_in, p.nsplit, n);
  EMIP_CHECK_LAUNCH("sum_parts");
  mem_bwd_m_kernel<<<B * nmb, NT, SMEM_BM, st>>>(p);
  EMIP_CHECK_LAUNCH("mem_bwd_m");
  return EMIP_OK;
}
