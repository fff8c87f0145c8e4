"""gym_rs::core -- Env / EnvProperties / ActionReward / RewardRange over the C ABI.

Reference: src/core.rs:25-57 (Env), :60-90 (EnvProperties), :94-106 (ActionReward),
:109-122 (RewardRange).

`ActionReward.reward: O64` and `.done: bool` are scalars in the reference (F9 in SURVEY.md), so
one class serves both shapes here:

  * num_envs == 1 and a Python int / float action  -> the reference's scalar surface:
    `step(action) -> ActionReward(observation=<Observation>, reward=float, done=bool, ...)`.
  * a device tensor of num_envs actions -> the batched surface: the fields of the returned
    ActionReward are zero-copy torch views of the handle's SoA device buffers
    (observation [obs_dim, num_envs] f32, reward [num_envs] f32, done / truncated [num_envs] u8).

Everything is computed by the CUDA kernels behind include/gymrs_b200.h.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Any, Generic, Optional, TypeVar

from . import _capi
from .spaces import BoxR, Discrete
from .utils.renderer import RenderMode, Renders

T = TypeVar("T")
E = TypeVar("E")


@dataclass
class ActionReward(Generic[T, E]):
    """src/core.rs:94-106"""
    observation: T
    reward: Any
    done: Any
    truncated: Any
    info: Optional[E]


@dataclass(frozen=True)
class RewardRange:
    """src/core.rs:109-122; default (-inf, +inf), :16-19"""
    lower_bound: float = -math.inf
    upper_bound: float = math.inf


@dataclass(frozen=True)
class Metadata:
    """src/utils/custom/structs.rs:13-19 (render metadata only; rendering is out of scope)"""
    render_modes: tuple
    render_fps: int


class _DevView:
    """A device buffer exposed through __cuda_array_interface__ (zero-copy torch.as_tensor)."""

    def __init__(self, ptr, shape, typestr, strides=None):
        self.__cuda_array_interface__ = {
            "shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
            "version": 2, "strides": strides,
        }


def _as_tensor(ptr, shape, typestr, strides, device):
    import torch
    return torch.as_tensor(_DevView(ptr, shape, typestr, strides), device=f"cuda:{device}")


class EnvProperties:
    """src/core.rs:60-90"""

    def metadata(self):
        return self._metadata

    def rand_random(self):
        """The generator that seeds states: (algorithm, key).  Reference returns &Pcg64."""
        return ("philox4x32_10", self._seed_used)

    def render_mode(self):
        return RenderMode.NONE  # core.rs:22,76-78

    def reward_range(self):
        lo, hi = C.c_double(), C.c_double()
        _capi.check(self._L.gymrs_reward_range(self._h, C.byref(lo), C.byref(hi)))
        return RewardRange(lo.value, hi.value)

    def action_space(self):
        return self._action_space

    def observation_space(self):
        return self._observation_space


class Env(EnvProperties):
    """src/core.rs:25-57.  Subclasses set KIND, OBSERVATION (a dataclass type), ACTION_DTYPE."""

    KIND: int = -1
    OBSERVATION = None
    STATE = None
    ACTION_DTYPE = "int32"
    INFO_ON_STEP = None        # cartpole: Some(()) -> (); mountain car: None
    INVALID_FMT = "{} usize invalid"
    WARNS_AFTER_TERMINATION = False  # CartPole logs a warning when stepped after termination (cartpole.rs:461)
    _METADATA = Metadata((), 0)

    def __init__(self, render_mode: RenderMode = RenderMode.NONE, num_envs: int = 1, device: int = 0,
                 global_env_offset: int = 0, params=None, time_limit: bool = False):
        if render_mode != RenderMode.NONE:
            # the reference asserts on unsupported modes (cartpole.rs:162); rendering is out of scope
            raise ValueError("only RenderMode.NONE is supported (SDL2 rendering is out of scope)")
        self._L = _capi.load()
        self._h = C.c_void_p()
        self.num_envs = int(num_envs)
        self.device = int(device)
        self._metadata = self._METADATA
        self._seed_used = None
        flags = _capi.FLAG_TIME_LIMIT if time_limit else 0
        pp = C.byref(params) if params is not None else None
        _capi.check(self._L.gymrs_create(self.KIND, self.num_envs, self.device, int(global_env_offset),
                                         pp, flags, C.byref(self._h)))
        self._refresh_views()
        self._refresh_spaces()
        # The device views above are torch tensors and callers produce actions with torch, so the
        # handle must run in torch's stream order, not on its private non-blocking stream.
        self._stream_seen = None
        self._follow_stream()

    # ---- plumbing ------------------------------------------------------------------
    def _follow_stream(self):
        """Keep the handle on torch's CURRENT stream of its device: step / rollout / reset call this
        first, so `with torch.cuda.stream(s):` and `torch.cuda.graph(...)` blocks order (or capture)
        the env's kernels like any torch op.  Switching is asynchronous (an event dependency from
        the old stream to the new one, gymrs_set_stream)."""
        cur = _current_raw_stream(self.device)
        if cur != self._stream_seen:
            self.set_stream(cur)
    def _refresh_views(self):
        b = _capi.Buffers()
        _capi.check(self._L.gymrs_get_buffers(self._h, C.byref(b)))
        self._buf = b
        n, ld, dev = int(b.num_envs), int(b.ld), self.device
        self.state_dim, self.obs_dim = int(b.state_dim), int(b.obs_dim)
        self._t_state = _as_tensor(b.state, (self.state_dim, n), "<f4", (4 * ld, 4), dev)
        self._t_obs = _as_tensor(b.obs, (self.obs_dim, n), "<f4", (4 * ld, 4), dev)
        self._t_reward = _as_tensor(b.reward, (n,), "<f4", None, dev)
        self._t_done = _as_tensor(b.done, (n,), "|u1", None, dev)
        self._t_truncated = _as_tensor(b.truncated, (n,), "|u1", None, dev)
        self._t_sbt = (_as_tensor(b.steps_beyond_terminated, (n,), "<i4", None, dev)
                       if b.steps_beyond_terminated else None)
        self._batched_result = ActionReward(self._t_obs, self._t_reward, self._t_done, self._t_truncated,
                                            self.INFO_ON_STEP)
        self._checked_action = None

    def _refresh_spaces(self):
        n = C.c_uint64()
        lo, hi = C.c_float(), C.c_float()
        _capi.check(self._L.gymrs_action_space(self._h, C.byref(n), C.byref(lo), C.byref(hi)))
        self._action_space = Discrete(int(n.value)) if n.value else BoxR(lo.value, hi.value)
        olo = (C.c_double * self.obs_dim)()
        ohi = (C.c_double * self.obs_dim)()
        _capi.check(self._L.gymrs_observation_space(self._h, olo, ohi))
        self._observation_space = BoxR(self.OBSERVATION(*olo), self.OBSERVATION(*ohi))

    @property
    def handle(self):
        return self._h

    @property
    def params(self):
        p = _capi.PARAMS[self.KIND]()
        _capi.check(self._L.gymrs_get_params(self._h, C.byref(p)))
        return p

    @params.setter
    def params(self, p):
        """The reference's physics constants are `pub` fields; assign a modified block back."""
        _capi.check(self._L.gymrs_set_params(self._h, C.byref(p)))
        self._refresh_spaces()

    def set_stream(self, cuda_stream: Optional[int]):
        """cuda_stream: a cudaStream_t as an integer (torch: stream.cuda_stream).  0 means CUDA's
        legacy default stream (what torch uses unless told otherwise) and is passed as the
        cudaStreamLegacy handle; None restores the handle's private stream."""
        if cuda_stream is None:
            ptr = 0
        else:
            ptr = int(cuda_stream) or 0x1  # cudaStreamLegacy
        _capi.check(self._L.gymrs_set_stream(self._h, C.c_void_p(ptr)))
        # a caller-chosen stream is followed until torch's current stream changes again
        self._stream_seen = None if cuda_stream is None else _current_raw_stream(self.device)

    def set_launch_config(self, vec: int = 0, block: int = 0, pdl: int = 1):
        _capi.check(self._L.gymrs_set_launch_config(self._h, vec, block, pdl))

    def set_launch_occupancy(self, wide: bool = False):
        """Step kernel built for a tighter register budget (more resident CTAs); see gymrs_b200.h."""
        _capi.check(self._L.gymrs_set_launch_occupancy(self._h, 1 if wide else 0))

    def sync(self):
        """Wait for queued work; raises like the reference's assert! if an action was invalid."""
        bad = C.c_uint64()
        rc = self._L.gymrs_sync(self._h, C.byref(bad))
        if rc == _capi.ERR_INVALID_ACTION:
            raise AssertionError(self._L.gymrs_last_error().decode())
        _capi.check(rc)

    # ---- `pub state` -----------------------------------------------------------------
    @property
    def state(self):
        """num_envs == 1: the STATE dataclass (like the reference's `pub state`); else the
        [state_dim, num_envs] device view."""
        if self.num_envs == 1:
            import numpy as np
            st = np.zeros((self.state_dim, 1), dtype=np.float32)
            _capi.check(self._L.gymrs_get_state(self._h, st.ctypes.data_as(C.c_void_p), None))
            return self.STATE(*[float(v) for v in st[:, 0]])
        self.sync()
        return self._t_state

    def get_state(self, with_sbt: bool = False):
        import numpy as np
        st = np.zeros((self.state_dim, self.num_envs), dtype=np.float32)
        sbt = np.zeros(self.num_envs, dtype=np.int32) if with_sbt else None
        _capi.check(self._L.gymrs_get_state(self._h, st.ctypes.data_as(C.c_void_p),
                                            sbt.ctypes.data_as(C.c_void_p) if with_sbt else None))
        return (st, sbt) if with_sbt else st

    def set_state(self, state, sbt=None):
        import numpy as np
        st = np.ascontiguousarray(np.asarray(state, dtype=np.float32).reshape(self.state_dim, self.num_envs))
        sb = None if sbt is None else np.ascontiguousarray(np.asarray(sbt, dtype=np.int32))
        _capi.check(self._L.gymrs_set_state(self._h, st.ctypes.data_as(C.c_void_p),
                                            None if sb is None else sb.ctypes.data_as(C.c_void_p)))

    # ---- Env ---------------------------------------------------------------------------
    def _scalar_action_array(self, action):
        import numpy as np
        if self.ACTION_DTYPE == "int32":
            if not isinstance(action, (int, np.integer)) or isinstance(action, bool):
                raise TypeError("action must be an integer (reference: Action = usize)")
            if action < 0 or action > 0x7FFFFFFF:
                raise AssertionError(self.INVALID_FMT.format(action))
            return np.array([action], dtype=np.int32)
        return np.array([action], dtype=np.float32)

    def step(self, action, autoreset: bool = False) -> ActionReward:
        """core.rs:42.  Scalar action (num_envs == 1) -> scalar ActionReward, synchronous.
        Device tensor of num_envs actions -> batched ActionReward of device views, asynchronous
        (call sync() to surface an invalid action)."""
        flags = _capi.STEP_AUTORESET if autoreset else 0
        if not hasattr(action, "data_ptr"):
            if self.num_envs != 1:
                raise TypeError("a scalar action needs num_envs == 1; pass a device tensor of actions")
            import numpy as np
            act = self._scalar_action_array(action)
            obs = np.zeros((self.obs_dim, 1), dtype=np.float32)
            rew = np.zeros(1, dtype=np.float32)
            dn = np.zeros(1, dtype=np.uint8)
            tr = np.zeros(1, dtype=np.uint8)
            vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
            _capi.check(self._L.gymrs_step_host(self._h, vp(act), flags, vp(obs), vp(rew), vp(dn), vp(tr)))
            bad = C.c_uint64()
            rc = self._L.gymrs_sync(self._h, C.byref(bad))
            if rc == _capi.ERR_INVALID_ACTION:
                raise AssertionError(self.INVALID_FMT.format(action))  # cartpole.rs:402-406
            _capi.check(rc)
            if self.WARNS_AFTER_TERMINATION and dn[0] and rew[0] == 0.0:
                # reward 0 on a terminal step: the env had terminated before this call (cartpole.rs:455-464)
                import logging
                logging.getLogger("gym_rs").warning(
                    "Calling step after termination may result in undefined behaviour. Consider reseting.")
            return ActionReward(self.OBSERVATION(*[float(v) for v in obs[:, 0]]), float(rew[0]),
                                bool(dn[0]), bool(tr[0]), self.INFO_ON_STEP)
        # A 1 M-env step is ~6 us of GPU time, so the host side of this call matters: the checks
        # are skipped for a tensor object that already passed them, and the result record (views of
        # the handle's fixed buffers) is built once.
        if action is not self._checked_action:
            self._check_actions(action)
            self._checked_action = action
        if _current_raw_stream(self.device) != self._stream_seen:
            self._follow_stream()
        rc = self._L.gymrs_step(self._h, action.data_ptr(), flags)
        if rc:
            _capi.check(rc)
        return self._batched_result

    def _check_actions(self, action, steps: int = 1):
        import torch
        want = torch.int32 if self.ACTION_DTYPE == "int32" else torch.float32
        if action.dtype != want or not action.is_cuda or action.device.index != self.device:
            raise TypeError(f"actions must be a {want} tensor on cuda:{self.device}")
        if action.numel() != self.num_envs * steps or not action.is_contiguous():
            raise ValueError("actions must be contiguous with num_envs elements per step")

    def step_host(self, actions, obs, reward, done, truncated=None, autoreset: bool = False):
        """gymrs_step_host: numpy / pinned-torch HOST buffers in, results copied out; synchronous."""
        flags = _capi.STEP_AUTORESET if autoreset else 0

        def hp(a):
            if a is None:
                return None
            return C.c_void_p(a.data_ptr()) if hasattr(a, "data_ptr") else a.ctypes.data_as(C.c_void_p)
        _capi.check(self._L.gymrs_step_host(self._h, hp(actions), flags, hp(obs), hp(reward), hp(done),
                                            hp(truncated)))

    def step_host_async(self, actions, obs, reward, done, truncated=None, autoreset: bool = False) -> int:
        """gymrs_step_host_async: like step_host but returns a ticket at once; the host buffers must
        not be touched until host_wait(ticket).  Two steps may be in flight (double buffering)."""
        flags = _capi.STEP_AUTORESET if autoreset else 0

        def hp(a):
            if a is None:
                return None
            return C.c_void_p(a.data_ptr()) if hasattr(a, "data_ptr") else a.ctypes.data_as(C.c_void_p)
        ticket = C.c_uint64()
        _capi.check(self._L.gymrs_step_host_async(self._h, hp(actions), flags, hp(obs), hp(reward), hp(done),
                                                  hp(truncated), C.byref(ticket)))
        return int(ticket.value)

    def host_wait(self, ticket: int):
        _capi.check(self._L.gymrs_host_wait(self._h, int(ticket)))

    def rollout_host(self, actions, obs=None, reward=None, done=None, truncated=None, n_steps=None,
                     autoreset: bool = True, on_step=None, u8_actions: bool = False, packed_done: bool = False):
        """gymrs_rollout_host: the host loop of examples/cartpole.rs:15-30 over the whole batch, run
        inside the library.  actions: HOST [action_slots, num_envs]; obs [result_slots, obs_dim,
        num_envs], reward / done / truncated [result_slots, num_envs] (pinned torch or numpy; done /
        truncated are [result_slots, ceil(num_envs / 8)] bit rows with packed_done).  Step t reads
        action slot t % action_slots and fills result slot t % result_slots; on_step(t, slot) is
        called in step order once the slot is complete."""
        def hp(a):
            if a is None:
                return None
            return a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data
        slots = [int(a.shape[0]) for a in (obs, reward, done, truncated) if a is not None]
        if len(set(slots)) > 1:
            raise ValueError("result buffers must have the same number of slots")
        d = _capi.HostRolloutDesc()
        d.actions, d.obs, d.reward, d.done, d.truncated = hp(actions), hp(obs), hp(reward), hp(done), hp(truncated)
        d.action_slots = int(actions.shape[0])
        d.result_slots = slots[0] if slots else 2
        d.transport = (_capi.HOST_U8_ACTIONS if u8_actions else 0) | (_capi.HOST_PACKED_DONE if packed_done else 0)
        cb = _capi.HOST_STEP_FN(lambda _user, t, slot: on_step(int(t), int(slot))) if on_step else _capi.HOST_STEP_FN()
        d.on_step = cb
        flags = _capi.STEP_AUTORESET if autoreset else 0
        self._follow_stream()
        _capi.check(self._L.gymrs_rollout_host(self._h, int(n_steps if n_steps is not None else d.action_slots),
                                               flags, C.byref(d)))

    def rollout(self, actions, obs_out=None, reward_out=None, done_out=None, autoreset: bool = True):
        """gymrs_rollout: actions [n_steps, num_envs] on the device; fused multi-step launch."""
        n_steps = int(actions.shape[0])
        self._check_actions(action=actions, steps=n_steps)
        self._follow_stream()
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())  # noqa: E731
        flags = _capi.STEP_AUTORESET if autoreset else 0
        _capi.check(self._L.gymrs_rollout(self._h, p(actions), n_steps, flags, p(obs_out), p(reward_out),
                                          p(done_out)))
        return ActionReward(self._t_obs, self._t_reward, self._t_done, self._t_truncated, self.INFO_ON_STEP)

    def reset(self, seed: Optional[int] = None, return_info: bool = False, options: Optional[BoxR] = None,
              mask=None):
        """core.rs:45-50.  options = BoxR(low_state, high_state) overrides the sampling bounds for
        this call (cartpole.rs:351-365).  mask (device u8 tensor) resets a subset."""
        import numpy as np
        ps = None
        if seed is not None:
            s = C.c_uint64(int(seed))
            ps = C.byref(s)
        lo = hi = None
        if options is not None:
            lo = np.asarray(_obs_values(options.low), dtype=np.float32)
            hi = np.asarray(_obs_values(options.high), dtype=np.float32)
            if lo.size != self.state_dim or hi.size != self.state_dim:
                raise ValueError("options bounds must have state_dim entries")
        used = C.c_uint64()
        self._follow_stream()
        _capi.check(self._L.gymrs_reset(
            self._h, ps, None if lo is None else lo.ctypes.data_as(C.c_void_p),
            None if hi is None else hi.ctypes.data_as(C.c_void_p),
            None if mask is None else C.c_void_p(mask.data_ptr()), C.byref(used)))
        self._seed_used = int(used.value)
        info = () if return_info else None  # cartpole.rs:511-515
        if self.num_envs == 1:
            return self._scalar_observation(), info  # the Observation type, like step() (core.rs:45-50)
        return self._t_obs, info

    def _scalar_observation(self):
        if self.OBSERVATION is self.STATE:
            return self.state
        self.sync()
        return self.OBSERVATION(*[float(v) for v in self._t_obs[:, 0].cpu()])

    def render(self, mode: RenderMode = RenderMode.NONE):
        return Renders.NONE  # renderer.rs:52-61 under RenderMode::None

    def close(self):
        """core.rs:56 -- releases the device buffers (the reference closes its SDL screen)."""
        if self._h:
            self._L.gymrs_destroy(self._h)
            self._h = C.c_void_p()

    def clone(self):
        """`Env: Clone` (core.rs:25): deep copy of the device state."""
        other = object.__new__(type(self))
        other.__dict__.update(self.__dict__)
        other._h = C.c_void_p()
        _capi.check(self._L.gymrs_clone(self._h, C.byref(other._h)))
        other._refresh_views()
        other._stream_seen = None
        other._follow_stream()
        return other

    def serialize(self):
        """`Env: Serialize` (core.rs:25, cartpole.rs:51): parameters + state as plain data.
        The RNG is skipped, as in the reference (serde(skip_serializing), cartpole.rs:85-86)."""
        p = self.params
        st, sbt = self.get_state(with_sbt=self._t_sbt is not None) if self._t_sbt is not None \
            else (self.get_state(), None)
        return {"kind": self.KIND, "num_envs": self.num_envs,
                "params": {k: getattr(p, k) for k, _ in p._fields_ if not k.startswith("_")},
                "state": st.tolist(), "steps_beyond_terminated": None if sbt is None else sbt.tolist()}

    # ---- checkpoint / resume (exact: carries the reset stream's key and step counter) ------
    def checkpoint(self):
        """gymrs_checkpoint_save: the whole handle as one numpy uint8 blob.  Unlike serialize()
        (which skips the RNG like the reference's serde derive), a handle restored from the blob
        continues bit-identically, auto-resets included."""
        import numpy as np
        nbytes = C.c_size_t()
        _capi.check(self._L.gymrs_checkpoint_size(self._h, C.byref(nbytes)))
        buf = np.empty(nbytes.value, dtype=np.uint8)
        _capi.check(self._L.gymrs_checkpoint_save(self._h, buf.ctypes.data_as(C.c_void_p), buf.nbytes))
        return buf

    def restore(self, blob):
        """gymrs_checkpoint_load into this handle (same kind, num_envs and time_limit)."""
        import numpy as np
        buf = np.ascontiguousarray(np.frombuffer(blob, dtype=np.uint8))
        _capi.check(self._L.gymrs_checkpoint_load(self._h, buf.ctypes.data_as(C.c_void_p), buf.nbytes))
        self._refresh_spaces()
        self._seed_used = checkpoint_info(buf)["seed"]

    @classmethod
    def from_checkpoint(cls, blob, device: int = 0):
        """gymrs_checkpoint_create: a new handle on `device` built from the blob alone."""
        import numpy as np
        import torch
        buf = np.ascontiguousarray(np.frombuffer(blob, dtype=np.uint8))
        info = checkpoint_info(buf)
        if info["kind"] != cls.KIND:
            raise ValueError(f"checkpoint holds env kind {info['kind']}, not {cls.KIND}")
        self = object.__new__(cls)
        self._L = _capi.load()
        self._h = C.c_void_p()
        self.num_envs = info["num_envs"]
        self.device = int(device)
        self._metadata = cls._METADATA
        self._seed_used = info["seed"]
        _capi.check(self._L.gymrs_checkpoint_create(buf.ctypes.data_as(C.c_void_p), buf.nbytes, self.device,
                                                    C.byref(self._h)))
        self._refresh_views()
        self._refresh_spaces()
        self._stream_seen = None
        self._follow_stream()
        return self

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def checkpoint_info(blob) -> dict:
    """gymrs_checkpoint_info_of: validates a blob (magic, version, size, checksum) on the host and
    returns its header fields; raises GymrsError for anything that is not an intact checkpoint."""
    import numpy as np
    buf = np.ascontiguousarray(np.frombuffer(blob, dtype=np.uint8))
    info = _capi.CheckpointInfo()
    _capi.check(_capi.load().gymrs_checkpoint_info_of(buf.ctypes.data_as(C.c_void_p), buf.nbytes, C.byref(info)))
    return {k: int(getattr(info, k)) for k, _ in info._fields_}


def _current_raw_stream(device: int) -> int:
    """torch's current stream on `device` as the integer gymrs_set_stream takes (0 = legacy default)."""
    import torch
    try:
        return int(torch._C._cuda_getCurrentRawStream(device))
    except AttributeError:  # pragma: no cover - older torch
        return int(torch.cuda.current_stream(device).cuda_stream)


def _obs_values(o):
    if hasattr(o, "__dataclass_fields__"):
        return [getattr(o, k) for k in o.__dataclass_fields__]
    return list(o)
