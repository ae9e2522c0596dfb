// b2q_rpm.cu — device replay memory kernels (include/b2q_rpm.h): batched ring append and uniform minibatch gather.
#include <cuda_runtime.h>
#include <cstdint>
#include "../../include/b2q_rpm.h"

namespace {
__global__ void rpm_append_kernel(float* s_obs, float* s_act, float* s_rew, float* s_next, float* s_term, const float* obs, const float* act, const float* rew,
                                  const float* next_obs, const float* term, int n, int od, int ad, int pos, int cap) {
  int i = blockIdx.x, t = threadIdx.x;
  if (i >= n) return;
  size_t slot = (size_t)((pos + i) % cap);
  for (int k = t; k < od; k += blockDim.x) { s_obs[slot * od + k] = obs[(size_t)i * od + k]; s_next[slot * od + k] = next_obs[(size_t)i * od + k]; }
  for (int k = t; k < ad; k += blockDim.x) s_act[slot * ad + k] = act[(size_t)i * ad + k];
  if (t == 0) { s_rew[slot] = rew[i]; s_term[slot] = term[i]; }
}
__device__ __forceinline__ uint32_t mix(uint64_t x) {  // splitmix64 finaliser
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; x ^= x >> 31;
  return (uint32_t)(x >> 32);
}
__global__ void rpm_sample_kernel(const float* s_obs, const float* s_act, const float* s_rew, const float* s_next, const float* s_term, float* obs, float* act,
                                  float* rew, float* next_obs, float* term, int batch, int od, int ad, int size, uint64_t seed) {
  int i = blockIdx.x, t = threadIdx.x;
  if (i >= batch) return;
  size_t slot = (size_t)(((uint64_t)mix(seed * 0x100000001B3ull + (uint64_t)i) * (uint64_t)size) >> 32);   // uniform in [0,size)
  for (int k = t; k < od; k += blockDim.x) { obs[(size_t)i * od + k] = s_obs[slot * od + k]; next_obs[(size_t)i * od + k] = s_next[slot * od + k]; }
  for (int k = t; k < ad; k += blockDim.x) act[(size_t)i * ad + k] = s_act[slot * ad + k];
  if (t == 0) { rew[i] = s_rew[slot]; term[i] = s_term[slot]; }
}
// ---- device-side cursor: state = {ring position, fill level, sample counter} lives in device memory, so that append / sample can be captured
//      once in a CUDA graph (kernel arguments frozen) and replayed every iteration
__global__ void rpm_append_cursor_kernel(float* s_obs, float* s_act, float* s_rew, float* s_next, float* s_term, const float* obs, const float* act, const float* rew,
                                         const float* next_obs, const float* term, int n, int od, int ad, int cap, const long long* state) {
  int i = blockIdx.x, t = threadIdx.x;
  if (i >= n) return;
  size_t slot = (size_t)((state[0] + i) % cap);
  for (int k = t; k < od; k += blockDim.x) { s_obs[slot * od + k] = obs[(size_t)i * od + k]; s_next[slot * od + k] = next_obs[(size_t)i * od + k]; }
  for (int k = t; k < ad; k += blockDim.x) s_act[slot * ad + k] = act[(size_t)i * ad + k];
  if (t == 0) { s_rew[slot] = rew[i]; s_term[slot] = term[i]; }
}
__global__ void rpm_advance_kernel(long long* state, int n, int cap) { state[0] = (state[0] + n) % cap; state[1] = state[1] + n < cap ? state[1] + n : cap; }
__global__ void rpm_sample_cursor_kernel(const float* s_obs, const float* s_act, const float* s_rew, const float* s_next, const float* s_term, float* obs, float* act,
                                         float* rew, float* next_obs, float* term, int batch, int od, int ad, uint64_t seed, const long long* state) {
  int i = blockIdx.x, t = threadIdx.x;
  if (i >= batch) return;
  const uint64_t size = (uint64_t)state[1], sd = seed + (uint64_t)state[2];
  size_t slot = (size_t)(((uint64_t)mix(sd * 0x100000001B3ull + (uint64_t)i) * size) >> 32);
  for (int k = t; k < od; k += blockDim.x) { obs[(size_t)i * od + k] = s_obs[slot * od + k]; next_obs[(size_t)i * od + k] = s_next[slot * od + k]; }
  for (int k = t; k < ad; k += blockDim.x) act[(size_t)i * ad + k] = s_act[slot * ad + k];
  if (t == 0) { rew[i] = s_rew[slot]; term[i] = s_term[slot]; }
}
__global__ void rpm_count_kernel(long long* state) { state[2] += 1; }
}  // namespace

extern "C" {
int b2q_rpm_append(float* s_obs, float* s_act, float* s_rew, float* s_next, float* s_term, const float* obs, const float* act, const float* rew,
                   const float* next_obs, const float* term, const uint8_t* valid, int n, int od, int ad, int pos, int cap, void* stream) {
  (void)valid;
  if (!s_obs || !s_act || !s_rew || !s_next || !s_term || !obs || !act || !rew || !next_obs || !term || n < 1 || cap < n || pos < 0) return -1;
  rpm_append_kernel<<<n, 64, 0, (cudaStream_t)stream>>>(s_obs, s_act, s_rew, s_next, s_term, obs, act, rew, next_obs, term, n, od, ad, pos, cap);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
int b2q_rpm_sample(const float* s_obs, const float* s_act, const float* s_rew, const float* s_next, const float* s_term, float* obs, float* act, float* rew,
                   float* next_obs, float* term, int batch, int od, int ad, int size, uint64_t seed, void* stream) {
  if (!s_obs || !obs || batch < 1 || size < 1) return -1;
  rpm_sample_kernel<<<batch, 64, 0, (cudaStream_t)stream>>>(s_obs, s_act, s_rew, s_next, s_term, obs, act, rew, next_obs, term, batch, od, ad, size, seed);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
int b2q_rpm_append_cursor(float* s_obs, float* s_act, float* s_rew, float* s_next, float* s_term, const float* obs, const float* act, const float* rew,
                          const float* next_obs, const float* term, int n, int od, int ad, int cap, long long* state, void* stream) {
  if (!s_obs || !s_act || !s_rew || !s_next || !s_term || !obs || !act || !rew || !next_obs || !term || !state || n < 1 || cap < n) return -1;
  rpm_append_cursor_kernel<<<n, 64, 0, (cudaStream_t)stream>>>(s_obs, s_act, s_rew, s_next, s_term, obs, act, rew, next_obs, term, n, od, ad, cap, state);
  rpm_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state, n, cap);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
int b2q_rpm_sample_cursor(const float* s_obs, const float* s_act, const float* s_rew, const float* s_next, const float* s_term, float* obs, float* act, float* rew,
                          float* next_obs, float* term, int batch, int od, int ad, uint64_t seed, long long* state, void* stream) {
  if (!s_obs || !obs || !state || batch < 1) return -1;
  rpm_sample_cursor_kernel<<<batch, 64, 0, (cudaStream_t)stream>>>(s_obs, s_act, s_rew, s_next, s_term, obs, act, rew, next_obs, term, batch, od, ad, seed, state);
  rpm_count_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
}
