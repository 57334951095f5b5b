"""HostPipeline: stepping for HOST-resident consumers (a policy that lives in host memory, a Gym
VectorEnv client) with the PCIe transfers overlapped -- split-batch double buffering.

The reference returns every step's observation to the caller's (host) memory
(helpers/gridworld_gym_env.py:525-585).  With a million environments per GPU the step kernel takes
0.11 ms while the result crosses PCIe in 1-2 ms, so the end-to-end rate is a question of bytes and
overlap, not of the kernel:

* the batch is split into `parts` slices, each an independent environment object with its own CUDA
  stream, device output tensors and pinned host buffers.  `submit(k, actions)` enqueues the slice's
  action upload, its fused step kernel and the download of its result and returns immediately;
  `wait(k)` blocks until that slice's result is in host memory.  While the caller consumes slice k
  (computes its next actions), the other slices' kernels and copies are in flight -- the EnvPool /
  Sample-Factory pattern.  Every slice still sees its own actions for step t before it produces step
  t: nothing is stale, the results are exactly those of the unsplit batch (the Philox streams are
  keyed by the global environment index);
* the default observation is the uint8 ASCII board (the reference's `ascii_codes`), 1 byte per
  cell instead of the 4 of the value-mapped float32 board; `value_board(obs)` applies the
  reference's value mapping (safety_game_mo_base.py DEFAULT_VALUE_MAPPING / the game's own) on the
  host, lazily, for consumers that want the float observation.

torch is used for device / pinned memory, streams and events only.  No CPU fallback.
"""
import numpy as np
import torch

from .vector_env import _ptr


class HostPipeline(object):
    """`make_part(n, env_index_base)` builds one slice (VectorEnv, ClassicVectorEnv, FiremakerVectorEnv, ...: anything
    with `step_raw(actions_ptr)` and the tensors named in `returns`).  `action_shape` is the per-environment action
    shape (() for single-agent games, (A,) for the multi-agent ones)."""

    def __init__(self, make_part, num_envs, device, env_index_base=0, parts=2, returns=("board", "reward", "terminated"),
                 action_shape=()):
        self.device = torch.device(device)
        self.num_envs = int(num_envs)
        parts = max(1, min(int(parts), self.num_envs))
        per, rem = divmod(self.num_envs, parts)
        self.bounds = []
        lo = 0
        for k in range(parts):
            hi = lo + per + (1 if k < rem else 0)
            self.bounds.append((lo, hi))
            lo = hi
        self.returns = tuple(returns)
        self.envs, self.streams, self.events = [], [], []
        self.h_actions, self.d_actions, self.h_out, self.d_out = [], [], [], []
        for lo, hi in self.bounds:
            env = make_part(hi - lo, env_index_base + lo)
            self.envs.append(env)
            self.streams.append(torch.cuda.Stream(self.device))
            self.events.append(torch.cuda.Event())
            shape = (hi - lo,) + tuple(action_shape)
            self.h_actions.append(torch.zeros(shape, dtype=torch.int32, pin_memory=True))
            self.d_actions.append(torch.zeros(shape, dtype=torch.int32, device=self.device))
            d = [getattr(env, name) for name in self.returns]
            self.d_out.append(d)
            self.h_out.append([torch.zeros(t.shape, dtype=t.dtype, pin_memory=True) for t in d])
        self.spec = getattr(self.envs[0], "spec", None)          # a mixed classic batch has one spec per type (`specs`)
        torch.cuda.synchronize(self.device)            # construction / reset ran on the current stream
        self._pending = [False] * parts

    @property
    def parts(self):
        return len(self.envs)

    def close(self):
        for s in self.streams:
            s.synchronize()
        for e in self.envs:
            e.close()
        self.envs = []

    def host_bytes_per_step(self):
        """(host->device, device->host) bytes one full step of all slices moves."""
        h2d = sum(a.numel() * a.element_size() for a in self.h_actions)
        d2h = sum(t.numel() * t.element_size() for outs in self.h_out for t in outs)
        return h2d, d2h

    def submit(self, k, actions_host=None):
        """Enqueue slice k's step: upload `actions_host` (pinned int32; None = the slice's own pinned action buffer
        `h_actions[k]`, filled in place by the caller), run the kernel, download the result.  Returns at once."""
        if self._pending[k]:
            raise RuntimeError("slice %d has a step in flight: wait(%d) first" % (k, k))
        src = self.h_actions[k] if actions_host is None else actions_host
        env = self.envs[k]
        with torch.cuda.stream(self.streams[k]):
            self.d_actions[k].copy_(src, non_blocking=True)
            rc = env.step_raw(_ptr(self.d_actions[k]))
            if rc != 0:
                from . import _abi
                _abi.check(rc)
            for h, d in zip(self.h_out[k], self.d_out[k]):
                h.copy_(d, non_blocking=True)
            self.events[k].record(self.streams[k])
        self._pending[k] = True

    def wait(self, k):
        """Blocks until slice k's submitted step is in host memory; returns its pinned host tensors in `returns` order
        (reused by the next submit of the slice: consume or copy before resubmitting)."""
        if not self._pending[k]:
            raise RuntimeError("slice %d has no step in flight" % k)
        self.events[k].synchronize()
        self._pending[k] = False
        return self.h_out[k]

    def step(self, actions_host):
        """Convenience: one step of the whole batch from one host action array; returns the per-slice host tensors."""
        for k, (lo, hi) in enumerate(self.bounds):
            self.h_actions[k].copy_(actions_host[lo:hi])
            self.submit(k)
        return [self.wait(k) for k in range(self.parts)]

    def value_board(self, board_host):
        """The reference's float32 observation from the uint8 ASCII board, on the host (value_mapping[chr])."""
        if self.spec is None:
            raise ValueError("a mixed batch has one value mapping per type: map the slices yourself")
        lut = np.zeros(256, np.float32)
        for ch, v in self.spec.value_mapping.items():
            lut[ord(ch)] = v
        b = board_host.numpy() if isinstance(board_host, torch.Tensor) else np.asarray(board_host)
        return lut[b]


def bind_process_to_gpu_numa_node(device_index):
    """Pins the calling process to the CPU cores next to GPU `device_index` (sysfs `local_cpulist` of its PCI function),
    so that pinned host buffers allocated afterwards are first-touched on the GPU's own NUMA node -- on an 8-GPU box the
    8 ranks' PCIe streams then land in both sockets' memory instead of one.  Returns the core list or None if the
    topology cannot be read (nothing is changed then)."""
    import os
    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (dom, bus, dev)
        with open(path) as f:
            text = f.read().strip()
        cores = set()
        for piece in text.split(","):
            if "-" in piece:
                a, b = piece.split("-")
                cores.update(range(int(a), int(b) + 1))
            elif piece:
                cores.add(int(piece))
        allowed = os.sched_getaffinity(0)
        cores &= allowed
        if not cores:
            return None
        os.sched_setaffinity(0, cores)
        return sorted(cores)
    except Exception:
        return None
