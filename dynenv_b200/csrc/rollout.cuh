/*
 * Rollout buffers on the device: what models/storage.py keeps per rollout step from the results of env.step(), without the
 * host round trip of models/train.py:269-284 (rewards / dones as numpy, agentFinished rebuilt from info['Full State']).
 * Both kernels are streaming HBM work: one thread per (environment, agent) column, consecutive threads on consecutive
 * floats of the [rollout_size][nEnvs * nAgents] buffers.
 */
#pragma once
#include <cstdint>

#include "../../include/dynenv_b200.h"

namespace dynenv {
namespace rollout {

struct InfoLayout {
  int words;      /* floats per result record */
  int rewards;    /* F [agents] */
  int done;       /* I */
  int epr, eppr;  /* F [agents] */
  int epor;       /* F [agents], -1: the environment has no observation rewards (episode_o_r = 0) */
  int epg;        /* I [2] */
  int state;      /* F [agents][state_stride] info['Full State'][0] */
  int state_stride;
};

/* storage.insert (models/storage.py:133-168): rewards[step], agentFinished[step] (= the last column of every agent's row of
   info['Full State'][0], models/train.py:276), dones[step]; storage.log_episode_rewards (:292-308) as per-environment running
   sums: episodes, sum of the mean episode_r / episode_p_r / episode_o_r over the agents, sum of episode_g[0..1], min and max of
   the mean episode_r; and the last DYNENV_EPISODE_RING episodes' per-agent episode_r (the reference's deque, for the median). */
__global__ void __launch_bounds__(256) insert_kernel(const float* __restrict__ info, InfoLayout L, long long nEnvs, int A,
                                                      float* __restrict__ rewards, uint8_t* __restrict__ finished,
                                                      uint8_t* __restrict__ dones, float* __restrict__ acc) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nEnvs * A) return;
  const long long e = t / A;
  const int a = (int)(t - e * A);
  const float* rec = info + e * L.words;
  rewards[t] = rec[L.rewards + a];
  finished[t] = rec[L.state + a * L.state_stride + L.state_stride - 1] != 0.f;
  if (a != 0) return;
  const int done = __float_as_int(rec[L.done]) != 0;
  dones[e] = (uint8_t)done;
  if (!done || !acc) return;
  float r = 0.f, pr = 0.f, orw = 0.f;
  for (int k = 0; k < A; k++) {
    r += rec[L.epr + k];
    pr += rec[L.eppr + k];
    if (L.epor >= 0) orw += rec[L.epor + k];
  }
  const float inv = 1.f / (float)A;
  r *= inv, pr *= inv, orw *= inv;
  float* s = acc + e * DYNENV_EPISODE_ACC_WORDS;
  const bool first = s[0] == 0.f;
  { /* the ring of the last episodes' per-agent rewards (median of print_reward_stats) */
    float* ring = s + DYNENV_EPISODE_STAT_WORDS + ((int)s[0] % DYNENV_EPISODE_RING) * DYNENV_EPISODE_RING_AGENTS;
    for (int k = 0; k < DYNENV_EPISODE_RING_AGENTS; k++) ring[k] = k < A ? rec[L.epr + k] : 0.f;
  }
  s[0] += 1.f;
  s[1] += r;
  s[2] += pr;
  s[3] += orw;
  s[4] += (float)__float_as_int(rec[L.epg]);
  s[5] += (float)__float_as_int(rec[L.epg + 1]);
  s[6] = first ? r : fminf(s[6], r);
  s[7] = first ? r : fmaxf(s[7], r);
}

/* storage._discount_rewards (models/storage.py:170-205): R = final_value; for i = T-1 .. 0: R = rewards[i] + discount * R;
   out[i] = R.  The reference builds its `dones4players` mask from zeros (:194), i.e. it never cuts the chain at an episode end;
   with `dones` given (an extension), R restarts from the reward alone at a step that finished its environment's episode.
   Multiply and add are rounded separately like the reference's float32 tensor ops. */
__global__ void __launch_bounds__(256) discount_kernel(const float* __restrict__ rewards, const float* __restrict__ final_value,
                                                        const uint8_t* __restrict__ dones, int A, float* __restrict__ out, int T,
                                                        long long n, float discount) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const long long nEnvs = n / A;
  const long long e = c / A;
  float R = final_value[c];
#pragma unroll 4
  for (int i = T - 1; i >= 0; i--) {
    const float r = rewards[(long long)i * n + c];
    const bool cut = dones && dones[(long long)i * nEnvs + e];
    R = cut ? r : __fadd_rn(r, __fmul_rn(discount, R));
    out[(long long)i * n + c] = R;
  }
}

}  // namespace rollout
}  // namespace dynenv
