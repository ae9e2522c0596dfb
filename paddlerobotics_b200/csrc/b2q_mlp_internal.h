// b2q_mlp_internal.h — library-internal extension of the MLP forward used by the SAC trainer (b2q_sac.cu): the same
// fused tcgen05 kernel, additionally dumping the bf16 layer inputs it already holds in shared memory, in the two layouts
// the backward GEMMs consume ([batch x width] and [width x batch]).  Not part of the public C ABI.
#pragma once
#include <cuda_bf16.h>
#include "../../include/b2q_mlp.h"

struct B2QMlpSaves {
  __nv_bfloat16* x_rm;   // [M][64]            concatenated, zero-padded input
  __nv_bfloat16* x_t;    // [64][M]
  __nv_bfloat16* h1_rm;  // [nets][M][256]     relu(layer 1)
  __nv_bfloat16* h1_t;   // [nets][256][M]
  __nv_bfloat16* h2_rm;  // [nets][M][256]     relu(layer 2)
  __nv_bfloat16* h2_t;   // [nets][256][M]
};
extern "C" int b2q_mlp_forward_ex(B2QMlpHandle h, const float* in1, int in1_dim, const float* in2, int M, int mode, uint64_t seed, const float* eps,
                                  float* out, float* logp, float* raw, const B2QMlpSaves* saves, void* stream);
