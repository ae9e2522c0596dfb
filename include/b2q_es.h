/* b2q_es.h — C ABI of the ES population-fitness kernels (K4).  Device pointers, caller's stream, 0 on success.
 *
 * Reference interfaces replaced (QuadrupedalRobots/ETGRL):
 *   b2q_es_accumulate  episode_reward += reward ... until done        train.py:213-249 (run_EStrain_episode)
 *   b2q_es_fitness     fitness_list.append(episode_reward)            train.py:404-413;
 *                      rewards gathered per individual                Dynamic_parallel_model.py:157-167
 * Env e = individual*rollouts + r.  The cross-GPU gather of `fitness` is the caller's NCCL all-gather.
 */
#ifndef B2Q_ES_H
#define B2Q_ES_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* per control step: for alive envs ret += reward, len += 1, alive &= !done.  elem_size 4 (float) or 8 (double). */
int b2q_es_accumulate(const void* reward, const uint8_t* done, uint8_t* alive, void* ret, int32_t* len, int n, int elem_size, void* stream);
/* fitness[i] = mean over the individual's rollouts of ret; mean_len (may be NULL) likewise for episode lengths. */
int b2q_es_fitness(const void* ret, const int32_t* len, void* fitness, void* mean_len, int pop, int rollouts, int elem_size, void* stream);
/* Batched ETG fit: for each individual i, points = prior_points + solutions[i].reshape(6,2) and
 * (w,b) = Opt_with_points(points=points, w0=w0, b0=b0) -- train.py:81-110,405-407 -- in float64 on the device.
 * obs6x20 = ETG_layer.update(t) at ts = [0.5T+0.1, 0, 0.05, 0.1, 0.15, 0.2] (row-major 6x20). w_out [pop,3,20], b_out [pop,3]. */
int b2q_etg_fit(const double* obs6x20, const double* prior_points, const double* solutions, const double* w0, const double* b0, double lamb,
                double precision, double* w_out, double* b_out, int pop, void* stream);
/* Dynamics identification: loss_func of model/Dynamic_parallel_model.py:29-41 accumulated on the device.
 * info [n,56] is the step kernel's info output; mean15/std15 = recorded {motor12, drpy3} statistics of THIS control step;
 * acc [n,15] running sums (zero it before an episode); reward[n] = 30 - (max_j mean_t motor_j + max_k mean_t drpy_k)/2. */
int b2q_dyn_accumulate(const void* info, const void* mean15, const void* std15, void* acc, int n, int elem_size, void* stream);
int b2q_dyn_finish(const void* acc, int steps, void* reward, int n, int elem_size, void* stream);
#ifdef __cplusplus
}
#endif
#endif
