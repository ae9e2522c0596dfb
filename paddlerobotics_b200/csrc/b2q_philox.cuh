// b2q_philox.cuh — counter-based N(0,1) draws shared by the MLP forward's rsample() epilogue (b2q_mlp.cu) and the actor-loss backward
// (b2q_sac.cu), which must see the SAME draw for element (row, col): Philox-4x32-10 keyed by a 64-bit seed, counter = (row, col), Box-Muller.
#pragma once
#include <cstdint>

namespace b2q_philox {

__device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0, uint32_t k1) {
  uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0, h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
  uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
  c0 = n0; c1 = l1; c2 = n2; c3 = l0;
}
__device__ __forceinline__ float philox_normal(uint64_t seed, uint32_t row, uint32_t col) {
  uint32_t c0 = row, c1 = col, c2 = 0x9E3779B9u, c3 = 0, k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int i = 0; i < 10; i++) { philox_round(c0, c1, c2, c3, k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
  float u1 = ((float)(c0 >> 8) + 0.5f) * (1.0f / 16777216.0f), u2 = ((float)(c1 >> 8) + 0.5f) * (1.0f / 16777216.0f);
  return sqrtf(-2.0f * logf(u1)) * cosf(6.283185307179586f * u2);
}
// Effective key of a launch: the host-side seed, advanced by a DEVICE-side counter when one is given (the learner's step counter), so that a
// CUDA-graph replay — whose kernel arguments are frozen at capture — still draws fresh noise every step.
__device__ __forceinline__ uint64_t effective_seed(uint64_t seed, const int* ctr) {
  return ctr ? seed + (uint64_t)(uint32_t)(*ctr) * 0x9E3779B97F4A7C15ull : seed;
}

}  // namespace b2q_philox
