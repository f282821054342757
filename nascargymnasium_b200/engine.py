"""ctypes binding of ``libncg_b200.so`` (the C ABI in ``include/ncg_b200.h``) exchanging PyTorch tensors.

PyTorch is plumbing here: it owns device memory and streams; every number is produced by the CUDA
library.  There is no CPU fallback: a missing library or a missing GPU raises."""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys
from typing import Optional, Sequence

import numpy as np

from . import layout as L
from . import track as T

_PKG = os.path.dirname(os.path.abspath(__file__))
# NCG_CHECKED=1: the range-checked build of the same source (csrc/ncg_defs.cuh NCG_CHECK), a debugging aid, several times slower
_CHECKED = os.environ.get("NCG_CHECKED", "") not in ("", "0")
# NCG_VARIANT=name NCG_DEFINES="-DX -DY": an A/B build of the same source under another file name (measurement only)
_VARIANT = os.environ.get("NCG_VARIANT", "").strip()
_SO = os.path.join(_PKG, f"libncg_b200_{_VARIANT}.so" if _VARIANT else ("libncg_b200_checked.so" if _CHECKED else "libncg_b200.so"))
_CSRC = os.path.join(_PKG, "csrc")
_SOURCES = ("ncg_b200.cu", "ncg_b200_cc.cu", "ncg_b200_res.cu", "ncg_step.cuh", "ncg_car.cuh", "ncg_b2.cuh", "ncg_defs.cuh")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-ldl"]


# Which Box2D 2.3.x b2CollidePolygons the contact path follows (include/ncg_b200.h NcgConfig.contacts): box2d-py 2.3.8
# bundles one of them and the two give different manifolds for ~2.6 % of touching box poses (profiles/r02_b2_version_study.json).
# NCG_BOX2D=2.3.0 selects the older form for every engine created with contacts=True.
DEFAULT_CONTACTS = 2 if os.environ.get("NCG_BOX2D", "").strip() == "2.3.0" else 1


class NcgError(RuntimeError):
    pass


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    srcs = [os.path.join(_CSRC, s) for s in _SOURCES] + [os.path.join(_PKG, "..", "include", "ncg_b200.h")]
    have_src = all(os.path.exists(s) for s in srcs)
    if os.path.exists(_SO) and not force:
        if not have_src or all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs):
            return _SO
    if not have_src:
        raise NcgError("CUDA sources missing and no prebuilt libncg_b200.so")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-DNCG_CHECKED"] if _CHECKED else []) + (os.environ.get("NCG_DEFINES", "").split() if _VARIANT else []) + (["-Xptxas", "-v"] if verbose else []) + \
          ["--threads", "3", "-o", _SO + ".tmp", os.path.join(_CSRC, "ncg_b200.cu"), os.path.join(_CSRC, "ncg_b200_cc.cu"), os.path.join(_CSRC, "ncg_b200_res.cu")]
    subprocess.check_call(cmd)
    os.replace(_SO + ".tmp", _SO)
    return _SO


class _Config(ctypes.Structure):
    _fields_ = [("device", ctypes.c_int32), ("num_envs", ctypes.c_int32), ("cars_per_env", ctypes.c_int32),
                ("discrete", ctypes.c_int32), ("reset_on_lap", ctypes.c_int32), ("auto_reset", ctypes.c_int32),
                ("contacts", ctypes.c_int32), ("track_info", ctypes.c_int32),
                ("start_x", ctypes.c_float), ("start_y", ctypes.c_float), ("start_angle", ctypes.c_float),
                ("car_contacts", ctypes.c_int32), ("grid_dx", ctypes.c_float), ("grid_dy", ctypes.c_float)]


class MappedBuffers(ctypes.Structure):
    """NcgMappedBuffers of include/ncg_b200.h: the buffers of one ncg_step_mapped_post."""
    _fields_ = [(k, ctypes.c_void_p) for k in ("actions", "obs", "reward", "terminated", "truncated", "final_obs", "ep_return", "ep_length")]


class Stats(ctypes.Structure):
    _fields_ = [("car_steps", ctypes.c_uint64), ("episodes", ctypes.c_uint64), ("laps", ctypes.c_uint64),
                ("ray_tests", ctypes.c_uint64), ("contact_steps", ctypes.c_uint64), ("toi_events", ctypes.c_uint64),
                ("overflow", ctypes.c_uint64), ("return_sum", ctypes.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


EXPORTS = ("ncg_last_error", "ncg_version", "ncg_create", "ncg_destroy", "ncg_upload_tracks", "ncg_reset", "ncg_step",
           "ncg_rollout", "ncg_step_host", "ncg_reset_host", "ncg_get_state", "ncg_set_state", "ncg_get_state_host",
           "ncg_set_state_host", "ncg_read_stats", "ncg_launch_count", "ncg_host_buffers", "ncg_step_pinned", "ncg_host_alloc",
           "ncg_host_free", "ncg_step_mapped", "ncg_plan_ctas", "ncg_set_rollout_base", "ncg_set_episode_outputs",
           "ncg_get_velocity_history_host", "ncg_set_track_redraw", "ncg_get_env_tracks", "ncg_get_car_pairs_host",
           "ncg_set_car_pairs_host", "ncg_debug_resident", "ncg_resident_pause", "ncg_step_mapped_from", "ncg_step_mapped_post", "ncg_step_mapped_wait")

_lib = None


def load_library():
    """dlopen the CUDA library (built in-tree).  Raises NcgError when it is missing: no fallback exists."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise NcgError(f"{_SO} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(nvcc, sm_100a). nascargymnasium_b200 has no CPU fallback.")
    lib = ctypes.CDLL(_SO)
    vp, i32, u64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_uint64
    lib.ncg_last_error.restype = ctypes.c_char_p
    lib.ncg_create.argtypes = [ctypes.POINTER(_Config), ctypes.POINTER(vp)]
    lib.ncg_destroy.argtypes = [vp]
    lib.ncg_upload_tracks.argtypes = [vp, vp, vp, i32]
    lib.ncg_reset.argtypes = [vp, vp, vp, i32, vp, vp]
    lib.ncg_step.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
    lib.ncg_rollout.argtypes = [vp, i32, u64, i32, vp, vp, vp, vp, vp]
    lib.ncg_step_host.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.ncg_reset_host.argtypes = [vp, vp, vp, i32, vp]
    lib.ncg_get_state.argtypes = [vp, vp, vp]
    lib.ncg_set_state.argtypes = [vp, vp, vp]
    lib.ncg_get_state_host.argtypes = [vp, vp]
    lib.ncg_set_state_host.argtypes = [vp, vp]
    lib.ncg_read_stats.argtypes = [vp, ctypes.POINTER(Stats), i32]
    lib.ncg_launch_count.argtypes = [vp]
    lib.ncg_host_buffers.argtypes = [vp] + [ctypes.POINTER(vp)] * 6
    lib.ncg_step_pinned.argtypes = [vp, i32, ctypes.POINTER(i32)]
    lib.ncg_launch_count.restype = ctypes.c_int64
    lib.ncg_host_alloc.argtypes = [ctypes.c_size_t, ctypes.POINTER(vp)]
    lib.ncg_host_free.argtypes = [vp]
    lib.ncg_step_mapped.argtypes = [vp] * 9 + [ctypes.POINTER(i32)]
    lib.ncg_plan_ctas.argtypes = [vp, i32, i32, i32, vp, vp, i32]
    lib.ncg_plan_ctas.restype = i32
    lib.ncg_set_rollout_base.argtypes = [vp, ctypes.c_uint32, ctypes.c_uint32]
    lib.ncg_set_episode_outputs.argtypes = [vp, vp, vp, vp]
    lib.ncg_get_velocity_history_host.argtypes = [vp, vp]
    lib.ncg_get_car_pairs_host.argtypes = [vp, vp]
    lib.ncg_set_car_pairs_host.argtypes = [vp, vp]
    lib.ncg_set_track_redraw.argtypes = [vp, i32, u64]
    lib.ncg_get_env_tracks.argtypes = [vp, vp]
    lib.ncg_debug_resident.argtypes = [vp, vp]
    lib.ncg_resident_pause.argtypes = [vp]
    lib.ncg_step_mapped_from.argtypes = [vp, vp, i32] + [vp] * 8 + [ctypes.POINTER(i32)]
    lib.ncg_step_mapped_post.argtypes = [vp, vp, i32, vp]
    lib.ncg_step_mapped_wait.argtypes = [vp, vp]
    _lib = lib
    return lib


def _check(rc: int):
    if rc == 0:
        return
    msg = load_library().ncg_last_error().decode()
    if rc == -1:
        raise ValueError(msg)
    if rc == -3:
        raise RuntimeError(msg)
    raise NcgError(f"ncg error {rc}: {msg}")


def _np_ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def plan_ctas(env_track: np.ndarray, cars_per_env: int, num_sms: int = 148):
    """(first_env, num_envs) per CTA of the step kernel's launch plan for this env -> track map (ncg_plan_ctas)."""
    lib = load_library()
    tid = np.ascontiguousarray(env_track, dtype=np.int32)
    first = np.zeros(len(tid), dtype=np.int32)
    count = np.zeros(len(tid), dtype=np.int32)
    n = lib.ncg_plan_ctas(_np_ptr(tid), len(tid), int(cars_per_env), int(num_sms), _np_ptr(first), _np_ptr(count), len(tid))
    if n < 0:
        _check(n)
    return first[:n].copy(), count[:n].copy()


class _HostAllocation:
    """Owner of one ncg_host_alloc allocation; freed when the last array over it dies."""

    def __init__(self, nbytes: int):
        self._lib = load_library()
        p = ctypes.c_void_p()
        _check(self._lib.ncg_host_alloc(nbytes, ctypes.byref(p)))
        self.ptr = p

    def __del__(self):
        try:
            if self.ptr is not None and self.ptr.value:
                self._lib.ncg_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


class HostBlock:
    """A page-locked, device-mapped host allocation (ncg_host_alloc) carved into named numpy arrays.  The arrays are
    what the kernel reads and writes directly; `busy()` tells whether the caller still holds views of any of them.  The
    memory lives as long as the block or any view of its arrays."""

    def __init__(self, fields):
        sizes = [int(np.prod(shape)) * np.dtype(dt).itemsize for _, shape, dt in fields]
        offs, total = [], 0
        for sz in sizes:
            offs.append(total)
            total += (sz + 255) // 256 * 256
        self._alloc = _HostAllocation(max(total, 256))
        base = self._alloc.ptr.value
        # numpy collapses .base chains to the first array over a foreign buffer, so every view a caller holds keeps a
        # reference on the flat array made here: its reference count says whether the block may be overwritten
        self._flat, self.arrays, self.ptrs = {}, {}, {}
        for (name, shape, dt), off in zip(fields, offs):
            n = int(np.prod(shape))
            raw = (ctypes.c_char * (n * np.dtype(dt).itemsize)).from_address(base + off)
            raw._owner = self._alloc                 # views -> flat -> raw -> allocation
            self._flat[name] = np.frombuffer(raw, dtype=dt, count=n)
            self.arrays[name] = self._flat[name].reshape(shape)
            self.ptrs[name] = ctypes.c_void_p(base + off)
        self._base_refs = {k: sys.getrefcount(a) for k, a in self._flat.items()}

        self.views, self._view_refs = (), ()

    def cache_views(self, views) -> None:
        """Shaped views of the arrays that a binding hands out on every step (made once instead of per step); `busy()` then
        also looks at who else holds THEM."""
        self.views = tuple(views)
        self._base_refs = {k: sys.getrefcount(a) for k, a in self._flat.items()}
        self._view_refs = self._view_counts()

    def _view_counts(self):
        return [sys.getrefcount(v) for v in self.views]

    def busy(self) -> bool:
        if any(sys.getrefcount(a) > self._base_refs[k] for k, a in self._flat.items()):
            return True
        return self._view_counts() != self._view_refs


class Engine:
    """E environments x C cars stepped on one B200.  Thin, explicit wrapper over the C ABI."""

    def __init__(self, num_envs: int, cars_per_env: int = 1, tracks: Sequence[str] = ("nascar",), discrete: bool = False,
                 reset_on_lap: bool = False, auto_reset: bool = True, contacts: bool = True, device: int = 0,
                 track_info: bool = False, start_position=(0.0, 0.0), start_angle: float = 0.0, car_contacts: bool = False,
                 grid=(8.0, 3.0)):
        lib = load_library()
        self._lib = lib
        self.num_envs, self.cars_per_env = int(num_envs), int(cars_per_env)
        self.num_cars = self.num_envs * self.cars_per_env
        self.discrete, self.device = bool(discrete), int(device)
        cfg = _Config(device, num_envs, cars_per_env, int(discrete), int(reset_on_lap), int(auto_reset),
                      DEFAULT_CONTACTS if contacts is True else int(contacts), int(track_info),
                      float(start_position[0]), float(start_position[1]), float(start_angle),
                      int(bool(car_contacts)), float(grid[0]), float(grid[1]))
        self.track_info = bool(track_info)
        h = ctypes.c_void_p()
        _check(lib.ncg_create(ctypes.byref(cfg), ctypes.byref(h)))
        self._h = h
        self.track_names = [T.track_name_of(t) for t in tracks]
        self.tables = [T.get_track_table(t) for t in tracks]
        blobs = [t.blob for t in self.tables]
        offs = np.zeros(len(blobs) + 1, dtype=np.int64)
        offs[1:] = np.cumsum([len(b) for b in blobs])
        blob = np.ascontiguousarray(np.concatenate(blobs), dtype=np.float32)
        _check(lib.ncg_upload_tracks(h, _np_ptr(blob), _np_ptr(offs), len(blobs)))

    # ------------------------------------------------------------------ torch (device tensor) path
    def _torch(self):
        import torch
        return torch

    def _stream(self):
        torch = self._torch()
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _ptr(t):
        return None if t is None else ctypes.c_void_p(t.data_ptr())

    def _dev(self, t, dtype, numel, name):
        torch = self._torch()
        if t is None:
            return None
        if not t.is_cuda or t.device.index != self.device or t.dtype != dtype or not t.is_contiguous() or t.numel() != numel:
            raise ValueError(f"{name}: expected a contiguous {dtype} CUDA tensor with {numel} elements on cuda:{self.device}")
        return t

    def reset(self, obs=None, env_mask=None, track_id=None, fresh: bool = True):
        torch = self._torch()
        E, N = self.num_envs, self.num_cars
        self._dev(obs, torch.float32, N * 38, "obs")
        self._dev(env_mask, torch.uint8, E, "env_mask")
        self._dev(track_id, torch.int32, E, "track_id")
        _check(self._lib.ncg_reset(self._h, self._ptr(env_mask), self._ptr(track_id), int(fresh), self._ptr(obs), self._stream()))

    def step(self, actions, obs, reward, terminated, truncated, final_obs=None):
        torch = self._torch()
        E, N = self.num_envs, self.num_cars
        self._dev(actions, torch.int32 if self.discrete else torch.float32, N if self.discrete else N * 2, "actions")
        self._dev(obs, torch.float32, N * 38, "obs")
        self._dev(reward, torch.float32, N, "reward")
        self._dev(terminated, torch.uint8, E, "terminated")
        self._dev(truncated, torch.uint8, E, "truncated")
        self._dev(final_obs, torch.float32, N * 38, "final_obs")
        _check(self._lib.ncg_step(self._h, self._ptr(actions), self._ptr(obs), self._ptr(reward), self._ptr(terminated),
                                  self._ptr(truncated), self._ptr(final_obs), self._stream()))

    def rollout(self, steps: int, seed: int = 0, mode: int = 0, obs_rollout=None, reward_rollout=None, done_rollout=None,
                obs_last=None):
        torch = self._torch()
        E, N = self.num_envs, self.num_cars
        self._dev(obs_rollout, torch.float32, steps * N * 38, "obs_rollout")
        self._dev(reward_rollout, torch.float32, steps * N, "reward_rollout")
        self._dev(done_rollout, torch.uint8, steps * E, "done_rollout")
        self._dev(obs_last, torch.float32, N * 38, "obs_last")
        _check(self._lib.ncg_rollout(self._h, steps, seed, mode, self._ptr(obs_rollout), self._ptr(reward_rollout),
                                     self._ptr(done_rollout), self._ptr(obs_last), self._stream()))

    def set_track_redraw(self, enable: bool, seed: int = 0):
        """Random-track mode: finished envs restart on another track (ncg_set_track_redraw)."""
        _check(self._lib.ncg_set_track_redraw(self._h, int(bool(enable)), int(seed) & 0xFFFFFFFFFFFFFFFF))

    def env_tracks(self) -> np.ndarray:
        out = np.empty(self.num_envs, dtype=np.int32)
        _check(self._lib.ncg_get_env_tracks(self._h, _np_ptr(out)))
        return out

    def set_rollout_base(self, car_base: int = 0, step_base: int = 0):
        """Philox counter offsets of rollout(): rank r of a sharded job passes car_base = r * num_cars."""
        _check(self._lib.ncg_set_rollout_base(self._h, int(car_base) & 0xFFFFFFFF, int(step_base) & 0xFFFFFFFF))

    def set_episode_outputs(self, ep_return=None, ep_length=None, any_done=None):
        """Device tensors step() fills for envs that finish: ep_return float32[N], ep_length int32[E], any_done int32[1]."""
        torch = self._torch()
        self._dev(ep_return, torch.float32, self.num_cars, "ep_return")
        self._dev(ep_length, torch.int32, self.num_envs, "ep_length")
        self._dev(any_done, torch.int32, 1, "any_done")
        self._ep_refs = (ep_return, ep_length, any_done)          # keep the storage alive while the engine points at it
        _check(self._lib.ncg_set_episode_outputs(self._h, self._ptr(ep_return), self._ptr(ep_length), self._ptr(any_done)))

    def get_state(self, out=None):
        torch = self._torch()
        if out is None:
            out = torch.empty((self.num_cars, L.RECORD_WORDS), dtype=torch.float32, device=f"cuda:{self.device}")
        self._dev(out, torch.float32, self.num_cars * L.RECORD_WORDS, "out")
        _check(self._lib.ncg_get_state(self._h, self._ptr(out), self._stream()))
        return out

    def set_state(self, records):
        torch = self._torch()
        self._dev(records, torch.float32, self.num_cars * L.RECORD_WORDS, "records")
        _check(self._lib.ncg_set_state(self._h, self._ptr(records), self._stream()))

    # ------------------------------------------------------------------ host (numpy) path
    def reset_host(self, env_mask: Optional[np.ndarray] = None, track_id: Optional[np.ndarray] = None, fresh: bool = True):
        obs = np.empty((self.num_cars, 38), dtype=np.float32)
        if env_mask is not None:
            env_mask = np.ascontiguousarray(env_mask, dtype=np.uint8)
        if track_id is not None:
            track_id = np.ascontiguousarray(track_id, dtype=np.int32)
        _check(self._lib.ncg_reset_host(self._h, _np_ptr(env_mask), _np_ptr(track_id), int(fresh), _np_ptr(obs)))
        return obs

    def step_host(self, actions: np.ndarray, want_final: bool = False):
        N, E = self.num_cars, self.num_envs
        actions = np.ascontiguousarray(actions, dtype=np.int32 if self.discrete else np.float32)
        if actions.size != (N if self.discrete else 2 * N):
            raise ValueError("actions: wrong number of elements")
        obs = np.empty((N, 38), dtype=np.float32)
        rew = np.empty(N, dtype=np.float32)
        te = np.empty(E, dtype=np.uint8)
        tr = np.empty(E, dtype=np.uint8)
        fin = np.empty((N, 38), dtype=np.float32) if want_final else None
        _check(self._lib.ncg_step_host(self._h, _np_ptr(actions), _np_ptr(obs), _np_ptr(rew), _np_ptr(te), _np_ptr(tr), _np_ptr(fin)))
        return obs, rew, te, tr, fin

    def pinned_views(self):
        """numpy views of the library's page-locked staging buffers (ncg_host_buffers)."""
        if getattr(self, "_views", None) is None:
            N, E = self.num_cars, self.num_envs
            ptrs = [ctypes.c_void_p() for _ in range(6)]
            _check(self._lib.ncg_host_buffers(self._h, *[ctypes.byref(p) for p in ptrs]))

            def view(ptr, ctype, n, shape):
                return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctype)), shape=(n,)).reshape(shape)
            act = view(ptrs[0], ctypes.c_int32, N, (N,)) if self.discrete else view(ptrs[0], ctypes.c_float, 2 * N, (N, 2))
            self._views = dict(actions=act, obs=view(ptrs[1], ctypes.c_float, N * 38, (N, 38)), reward=view(ptrs[2], ctypes.c_float, N, (N,)),
                               terminated=view(ptrs[3], ctypes.c_uint8, E, (E,)), truncated=view(ptrs[4], ctypes.c_uint8, E, (E,)),
                               final_obs=view(ptrs[5], ctypes.c_float, N * 38, (N, 38)))
        return self._views

    def step_pinned(self, want_final: bool = True) -> bool:
        """Step with actions already written into pinned_views()['actions']; results are read in place."""
        done = ctypes.c_int32(0)
        _check(self._lib.ncg_step_pinned(self._h, int(want_final), ctypes.byref(done)))
        return bool(done.value)

    def result_block(self) -> "HostBlock":
        """A fresh set of mapped host result buffers for step_mapped."""
        self._lib.ncg_resident_pause(self._h)        # (page-locked allocations synchronise the device: do not wait for a resident kernel's idle time)
        N, E = self.num_cars, self.num_envs
        return HostBlock([("obs", (N, 38), np.float32), ("reward", (N,), np.float32), ("terminated", (E,), np.uint8),
                          ("truncated", (E,), np.uint8)])

    def aux_block(self) -> "HostBlock":
        """Mapped host buffers for the actions and the rarely read per-episode outputs of step_mapped."""
        self._lib.ncg_resident_pause(self._h)
        N, E = self.num_cars, self.num_envs
        act = ("actions", (N,), np.int32) if self.discrete else ("actions", (N, 2), np.float32)
        return HostBlock([act, ("final_obs", (N, 38), np.float32), ("ep_return", (N,), np.float32), ("ep_length", (E,), np.int32)])

    def step_mapped(self, aux: "HostBlock", res: "HostBlock") -> bool:
        """One step: actions from aux['actions'], results into `res` (written by the kernel across PCIe).  Returns
        True when at least one env finished (then aux final_obs / ep_return / ep_length hold those envs' values)."""
        done = ctypes.c_int32(0)
        a, r = aux.ptrs, res.ptrs
        _check(self._lib.ncg_step_mapped(self._h, a["actions"], r["obs"], r["reward"], r["terminated"], r["truncated"],
                                         a["final_obs"], a["ep_return"], a["ep_length"], ctypes.byref(done)))
        return bool(done.value)

    def get_state_host(self) -> np.ndarray:
        out = np.empty((self.num_cars, L.RECORD_WORDS), dtype=np.float32)
        _check(self._lib.ncg_get_state_host(self._h, _np_ptr(out)))
        return out

    def set_state_host(self, records: np.ndarray) -> None:
        records = np.ascontiguousarray(records, dtype=np.float32)
        if records.size != self.num_cars * L.RECORD_WORDS:
            raise ValueError("records: wrong number of elements")
        _check(self._lib.ncg_set_state_host(self._h, _np_ptr(records)))

    def get_car_pairs_host(self) -> np.ndarray:
        """car_contacts only: the car-car contact tables of the shared worlds, (E, NCG_CAR_PAIR_WORDS) float32 (raw words)."""
        out = np.empty((self.num_envs, L.DEFS["NCG_CAR_PAIR_WORDS"]), dtype=np.float32)
        _check(self._lib.ncg_get_car_pairs_host(self._h, _np_ptr(out)))
        return out

    def set_car_pairs_host(self, pairs: np.ndarray) -> None:
        pairs = np.ascontiguousarray(pairs, dtype=np.float32)
        if pairs.size != self.num_envs * L.DEFS["NCG_CAR_PAIR_WORDS"]:
            raise ValueError("pairs: wrong number of elements")
        _check(self._lib.ncg_set_car_pairs_host(self._h, _np_ptr(pairs)))

    def velocity_history_host(self) -> np.ndarray:
        """(N, 600, 2) ring of the pre-step velocities of the running episode (track_info engines only)."""
        out = np.empty((self.num_cars, L.VEL_HISTORY, 2), dtype=np.float32)
        _check(self._lib.ncg_get_velocity_history_host(self._h, _np_ptr(out)))
        return out

    def read_stats(self, reset: bool = False) -> dict:
        s = Stats()
        _check(self._lib.ncg_read_stats(self._h, ctypes.byref(s), int(reset)))
        return s.as_dict()

    def resident_pause(self) -> None:
        """End a resident launch of the step kernel now (ncg_resident_pause): before CUDA calls that synchronise the device."""
        _check(self._lib.ncg_resident_pause(self._h))

    @property
    def resident_stats(self) -> dict:
        """Diagnostics of the resident step kernel behind step_mapped (ends a running resident launch)."""
        out = np.zeros(4, dtype=np.uint64)
        _check(self._lib.ncg_debug_resident(self._h, _np_ptr(out)))
        w, n, d, m = (int(x) for x in out)
        return {"steps": n, "host_us_per_step": w / max(n, 1) / 1e3, "device_us_per_step": d / max(m, 1) / 1e3, "device_steps": m}

    @property
    def launch_count(self) -> int:
        return int(self._lib.ncg_launch_count(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ncg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
