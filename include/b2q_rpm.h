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
#ifdef __cplusplus
}
#endif
#endif
