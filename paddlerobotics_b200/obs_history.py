"""Observation history stack for recurrent / stacked-observation policies — the batched, device-resident counterpart of
`ObservationWrapper` (deployment/envs/EnvWrapper.py:195-241; selected by sensor_mode["RNN"] = {time_steps, time_interval,
mode}, train.py:274-277).  Same semantics, including the reference's quirks: the history holds time_steps*time_interval past
observations, starts as zeros except the newest slot (= the reset observation), the stacked row is
[hist[0], hist[interval], ..., hist[(T-1)*interval], current], and the history shifts by one per call.

    h = ObservationHistory(num_envs, obs_dim, time_steps=5, time_interval=1, mode="stack")
    stacked = h.reset(obs)          # [N, (T+1)*obs_dim]   ("GRU": [N, T+1, obs_dim])
    stacked = h.push(next_obs, done_mask=done)   # rows of finished envs restart their history from next_obs
"""
import torch


class ObservationHistory:
    def __init__(self, num_envs, obs_dim, time_steps=5, time_interval=1, mode="stack", device="cuda", dtype=torch.float32):
        if mode not in ("stack", "GRU", "None"):
            raise ValueError("mode must be 'stack', 'GRU' or 'None'")
        self.n, self.d, self.T, self.I, self.mode = int(num_envs), int(obs_dim), int(time_steps), int(time_interval), mode
        self.hist = torch.zeros(self.T * self.I, self.n, self.d, device=device, dtype=dtype) if self.T > 0 else None
        self.obs_dim = self.d * (self.T + 1) if (self.T > 0 and mode == "stack") else self.d      # get_obs_dim, EnvWrapper.py:199-204

    def _out(self, obs):
        if self.T <= 0 or self.mode == "None":
            return obs
        rows = [self.hist[t * self.I] for t in range(self.T)] + [obs]
        st = torch.stack(rows, dim=1)                                 # [N, T+1, d]
        return st if self.mode == "GRU" else st.reshape(self.n, -1)

    def reset(self, obs, env_mask=None):
        """EnvWrapper.py:224-238.  env_mask (bool [N]) restarts only those rows (batched auto-reset)."""
        if self.T <= 0:
            return obs
        if env_mask is None:
            self.hist.zero_()
            out = self._out(obs)
            self.hist[-1] = obs
            return out
        m = env_mask.bool()
        self.hist[:, m] = 0
        out = self._out(obs)
        self.hist[-1, m] = obs[m]
        return out

    def push(self, obs, done_mask=None):
        """get_observation, EnvWrapper.py:206-222: stacked row from the history BEFORE this observation, then shift it in.
        Rows flagged in done_mask (their `obs` is the reset observation of the next episode) restart like reset()."""
        if self.T <= 0:
            return obs
        if done_mask is not None and bool(done_mask.any()):
            m = done_mask.bool()
            self.hist[:, m] = 0
            out = self._out(obs)
            self.hist[:-1, ~m] = self.hist[1:, ~m].clone()
            self.hist[-1] = obs
            return out
        out = self._out(obs)
        self.hist[:-1] = self.hist[1:].clone()
        self.hist[-1] = obs
        return out
