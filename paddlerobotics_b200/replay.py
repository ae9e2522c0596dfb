"""Device-resident replay memory with the call surface of parl.utils.ReplayMemory as the reference uses it
(`rpm.append(obs, action, reward, next_obs, terminal)`, `rpm.sample_batch(B)`, `rpm.size()`: ETGRL/train.py:159,164,142),
batched over envs: append() takes [N, ...] device tensors (one transition per env per control step)."""
import ctypes as C

import torch

from . import _lib


class ReplayMemory:
    def __init__(self, max_size, obs_dim, act_dim, device=0, device_cursor=False):
        """device_cursor=True keeps the ring position, the fill level and the sample counter in device memory (b2q_rpm_*_cursor): append /
        sample_batch then take nothing step-dependent as a kernel argument and can be captured in a CUDA graph.  The host-side mirrors
        (size()) advance identically."""
        self.lib = _lib.load()
        self.max_size, self.obs_dim, self.act_dim = int(max_size), obs_dim, act_dim
        self.device = torch.device("cuda", int(device))
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=self.device)
        self.obs, self.next_obs = z(self.max_size, obs_dim), z(self.max_size, obs_dim)
        self.action, self.reward, self.terminal = z(self.max_size, act_dim), z(self.max_size), z(self.max_size)
        self._curr_size, self._curr_pos, self._samples = 0, 0, 0
        self.cursor = torch.zeros(3, dtype=torch.int64, device=self.device) if device_cursor else None   # {position, fill level, samples}

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def size(self):
        return self._curr_size

    def __len__(self):
        return self._curr_size

    def append(self, obs, act, reward, next_obs, terminal):
        t = lambda x: torch.as_tensor(x, dtype=torch.float32, device=self.device).contiguous()
        obs, act, reward, next_obs, terminal = t(obs).reshape(-1, self.obs_dim), t(act).reshape(-1, self.act_dim), t(reward).reshape(-1), t(next_obs).reshape(-1, self.obs_dim), t(terminal).reshape(-1)
        n = obs.shape[0]
        if self.cursor is not None:
            rc = self.lib.b2q_rpm_append_cursor(self.obs.data_ptr(), self.action.data_ptr(), self.reward.data_ptr(), self.next_obs.data_ptr(), self.terminal.data_ptr(),
                                                obs.data_ptr(), act.data_ptr(), reward.data_ptr(), next_obs.data_ptr(), terminal.data_ptr(),
                                                n, self.obs_dim, self.act_dim, self.max_size, self.cursor.data_ptr(), self._stream())
            assert rc == 0, rc
            self._keep = (obs, act, reward, next_obs, terminal)      # inputs of a captured launch must outlive the capture
            self._curr_pos = (self._curr_pos + n) % self.max_size
            self._curr_size = min(self._curr_size + n, self.max_size)
            return
        rc = self.lib.b2q_rpm_append(self.obs.data_ptr(), self.action.data_ptr(), self.reward.data_ptr(), self.next_obs.data_ptr(), self.terminal.data_ptr(),
                                     obs.data_ptr(), act.data_ptr(), reward.data_ptr(), next_obs.data_ptr(), terminal.data_ptr(), None,
                                     n, self.obs_dim, self.act_dim, self._curr_pos, self.max_size, self._stream())
        assert rc == 0, rc
        self._curr_pos = (self._curr_pos + n) % self.max_size
        self._curr_size = min(self._curr_size + n, self.max_size)

    def advance(self, n, samples=1):
        """Host-side mirrors only: a captured iteration (device cursor) was replayed — n rows appended, `samples` minibatches drawn."""
        self._curr_pos = (self._curr_pos + n) % self.max_size
        self._curr_size = min(self._curr_size + n, self.max_size)
        self._samples += samples

    def sample_batch(self, batch_size, seed=None, out=None):
        """Uniform sample (replay_memory.py sample_batch).  out = (obs, act, rew, next_obs, term) float32 device tensors to gather into
        (e.g. SACLearner.static_batch(): the learner's CUDA-graph inputs are then filled in place, with no copy in between)."""
        self._samples += 1
        if out is not None:
            obs, act, rew, nobs, term = out
            assert obs.shape == (batch_size, self.obs_dim) and act.shape == (batch_size, self.act_dim) and all(x.is_contiguous() and x.dtype == torch.float32 for x in out)
        else:
            z = lambda *s: torch.empty(*s, dtype=torch.float32, device=self.device)
            obs, nobs, act, rew, term = z(batch_size, self.obs_dim), z(batch_size, self.obs_dim), z(batch_size, self.act_dim), z(batch_size), z(batch_size)
        if self.cursor is not None:
            rc = self.lib.b2q_rpm_sample_cursor(self.obs.data_ptr(), self.action.data_ptr(), self.reward.data_ptr(), self.next_obs.data_ptr(), self.terminal.data_ptr(),
                                                obs.data_ptr(), act.data_ptr(), rew.data_ptr(), nobs.data_ptr(), term.data_ptr(), batch_size, self.obs_dim, self.act_dim,
                                                C.c_uint64(0 if seed is None else seed), self.cursor.data_ptr(), self._stream())
            assert rc == 0, rc
            return obs, act, rew, nobs, term
        rc = self.lib.b2q_rpm_sample(self.obs.data_ptr(), self.action.data_ptr(), self.reward.data_ptr(), self.next_obs.data_ptr(), self.terminal.data_ptr(),
                                     obs.data_ptr(), act.data_ptr(), rew.data_ptr(), nobs.data_ptr(), term.data_ptr(), batch_size, self.obs_dim, self.act_dim,
                                     self._curr_size, C.c_uint64(self._samples if seed is None else seed), self._stream())
        assert rc == 0, rc
        return obs, act, rew, nobs, term
