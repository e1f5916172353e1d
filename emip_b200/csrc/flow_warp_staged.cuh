// K3 shared-memory-staged variant (flow_warp_staged.cu): internal launcher used by flow_warp.cu.
#pragma once
#include "common.cuh"

// Returns EMIP_OK after launching, EMIP_ENOSYS when the shape / alignment is not covered (the caller then uses the
// direct-gather kernel), or an error code.  bwd = false: out = warped image [B,3,H,W]; bwd = true: out = dflow [B,2,H,W].
// Stateless; dev_stats / host_stats / seq: optional caller-owned kernel-choice feedback (see flow_warp_staged.cu).
int flow_warp_staged_launch(bool bwd, const float* x, const float* flow, const float* dout, float* out, int B, int H, int W,
                            long long fsb, long long fsc, int ctas_per_sm, unsigned* dev_stats, unsigned* host_stats,
                            unsigned seq, cudaStream_t st);
