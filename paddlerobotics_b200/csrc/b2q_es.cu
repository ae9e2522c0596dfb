// b2q_es.cu — K4: ES population fitness on device (include/b2q_es.h).
// Replaces `fitness_list.append(episode_reward)` (train.py:404-413) and the per-actor reward gather of
// Dynamic_parallel_model.py:152-167: per-env episode return / length accumulation with first-done freezing, then a
// segmented mean over each individual's rollouts.  The all-gather across GPUs is done by the caller (NCCL).
#include <cuda_runtime.h>
#include <cstdint>
#include "../../include/b2q_es.h"

namespace {

template <typename T>
__global__ void es_accumulate_kernel(const T* __restrict__ reward, const uint8_t* __restrict__ done, uint8_t* __restrict__ alive,
                                     T* __restrict__ ret, int32_t* __restrict__ len, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (alive[i]) {               // run_EStrain_episode: accumulate until the env reports done (train.py:221-246)
    ret[i] += reward[i];
    len[i] += 1;
    if (done[i]) alive[i] = 0;
  }
}

// one warp per individual: deterministic shuffle-tree sum over its `rollouts` consecutive envs
template <typename T>
__global__ void es_fitness_kernel(const T* __restrict__ ret, const int32_t* __restrict__ len, T* __restrict__ fitness, T* __restrict__ mean_len,
                                  int pop, int rollouts) {
  int ind = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (ind >= pop) return;
  T s = 0, l = 0;
  for (int r = lane; r < rollouts; r += 32) { s += ret[(size_t)ind * rollouts + r]; l += (T)len[(size_t)ind * rollouts + r]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); l += __shfl_xor_sync(0xffffffffu, l, o); }
  if (lane == 0) { fitness[ind] = s / (T)rollouts; if (mean_len) mean_len[ind] = l / (T)rollouts; }
}

}  // namespace

extern "C" {

int b2q_es_accumulate(const void* reward, const uint8_t* done, uint8_t* alive, void* ret, int32_t* len, int n, int elem_size, void* stream) {
  if (!reward || !done || !alive || !ret || !len || n < 1 || (elem_size != 4 && elem_size != 8)) return -1;
  cudaStream_t s = (cudaStream_t)stream;
  if (elem_size == 4) es_accumulate_kernel<float><<<(n + 255) / 256, 256, 0, s>>>((const float*)reward, done, alive, (float*)ret, len, n);
  else es_accumulate_kernel<double><<<(n + 255) / 256, 256, 0, s>>>((const double*)reward, done, alive, (double*)ret, len, n);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int b2q_es_fitness(const void* ret, const int32_t* len, void* fitness, void* mean_len, int pop, int rollouts, int elem_size, void* stream) {
  if (!ret || !len || !fitness || pop < 1 || rollouts < 1 || (elem_size != 4 && elem_size != 8)) return -1;
  cudaStream_t s = (cudaStream_t)stream;
  int threads = 128, blocks = (pop * 32 + threads - 1) / threads;
  if (elem_size == 4) es_fitness_kernel<float><<<blocks, threads, 0, s>>>((const float*)ret, len, (float*)fitness, (float*)mean_len, pop, rollouts);
  else es_fitness_kernel<double><<<blocks, threads, 0, s>>>((const double*)ret, len, (double*)fitness, (double*)mean_len, pop, rollouts);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // extern "C"
