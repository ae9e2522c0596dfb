/* b2q_rpm.h — device-resident replay memory: the ring storage lives in caller-owned device tensors,
 * these two kernels append a whole batch of env transitions per control step and gather a uniformly sampled minibatch,
 * so the SAC loop has no host traffic.  Replaces parl.utils.ReplayMemory.append / sample_batch as used at
 * ETGRL/train.py:159,164 (and BCreplay_buffer.py:21-84).  float32 storage.  0 on success. */
#ifndef B2Q_RPM_H
#define B2Q_RPM_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* writes n rows at ring positions (pos + i) % capacity. */
int b2q_rpm_append(float* s_obs, float* s_act, float* s_rew, float* s_next, float* s_term,
                   const float* obs, const float* act, const float* rew, const float* next_obs, const float* term, const uint8_t* valid_or_null,
                   int n, int obs_dim, int act_dim, int pos, int capacity, void* stream);
/* gathers `batch` rows with indices drawn uniformly in [0,size) from a counter RNG keyed by seed. */
int b2q_rpm_sample(const float* s_obs, const float* s_act, const float* s_rew, const float* s_next, const float* s_term,
                   float* obs, float* act, float* rew, float* next_obs, float* term, int batch, int obs_dim, int act_dim, int size,
                   uint64_t seed, void* stream);
/* The same two operations with a DEVICE-side cursor: state = device int64[3] {ring position, fill level, sample counter}.  append writes at
 * (state[0] + i) % capacity and then advances position and fill level; sample draws from [0, state[1]) with the key seed + state[2] and then
 * increments the counter.  Nothing about the ring's progress is a kernel argument, so a whole training iteration (policy forward, env step,
 * append, sample, learn) can be captured ONCE in a CUDA graph and replayed.  The fill level must be >= 1 when sample runs. */
int b2q_rpm_append_cursor(float* s_obs, float* s_act, float* s_rew, float* s_next, float* s_term,
                          const float* obs, const float* act, const float* rew, const float* next_obs, const float* term,
                          int n, int obs_dim, int act_dim, int capacity, long long* state, void* stream);
int b2q_rpm_sample_cursor(const float* s_obs, const float* s_act, const float* s_rew, const float* s_next, const float* s_term,
                          float* obs, float* act, float* rew, float* next_obs, float* term, int batch, int obs_dim, int act_dim,
                          uint64_t seed, long long* state, void* stream);
#ifdef __cplusplus
}
#endif
#endif
