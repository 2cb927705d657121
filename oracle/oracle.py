"""ctypes binding of the CPU oracle (oracle/libppg_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this module (as the checker / the reported CPU baseline).  The product package
predpreygrass_b200 never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from predpreygrass_b200.config import N_STATS, PpgBuffers, PpgConfig, PpgTape

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libppg_oracle.so")


class OracleBuffers(C.Structure):
    _fields_ = [("f", PpgBuffers), ("obs64", C.c_void_p * 2), ("reward64", C.c_void_p * 2)]


def build(force=False):
    src = [os.path.join(HERE, f) for f in ("ppg_oracle.c", "ppg_oracle_eco.c", "ppg_oracle_stag.c", "ppg_oracle.h", "ppg_oracle_int.h", "Makefile")]
    src += [os.path.join(HERE, "..", "include", f) for f in ("ppg.h", "ppg_philox.h")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.ppgo_create.restype = C.c_void_p
        L.ppgo_create.argtypes = [C.POINTER(PpgConfig), C.c_int32]
        L.ppgo_destroy.argtypes = [C.c_void_p]
        L.ppgo_load_tape.argtypes = [C.c_void_p, C.POINTER(PpgTape)]
        L.ppgo_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ppgo_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ppgo_random_actions.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        L.ppgo_get_buffers.argtypes = [C.c_void_p, C.POINTER(OracleBuffers)]
        L.ppgo_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.ppgo_read_env.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 9
        L.ppgo_read_grid.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.ppgo_set_threads.argtypes = [C.c_void_p, C.c_int32]
        L.ppgo_env_step_ordered.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ppgo_env_reset_cells.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.ppgo_env_agents.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        L.ppgo_env_reset_eco.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        L.ppgo_read_env_eco.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 6
        L.ppgo_env_reset_trait.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        L.ppgo_read_env_acc.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        L.ppgo_read_episode_eco.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        L.ppgo_read_episode_events_eco.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.ppgo_env_reset_stag.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ppgo_read_env_stag.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 6
        _lib = L
    return _lib


def _np(ptr, shape, dtype):
    n = int(np.prod(shape))
    if n == 0 or not ptr:
        return np.zeros(shape, dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def _flatten(seqs, dtype):
    off = np.zeros(len(seqs) + 1, np.int64)
    for i, c in enumerate(seqs):
        off[i + 1] = off[i] + len(c)
    flat = np.concatenate([np.asarray(c, dtype) for c in seqs]) if off[-1] else np.zeros(1, dtype)
    return np.ascontiguousarray(flat, dtype), off


def make_tape(cells_per_env, reals_per_env=None):
    """cells_per_env / reals_per_env: lists of sequences -> (PpgTape, keepalive)"""
    flat, off = _flatten(cells_per_env, np.int32)
    t = PpgTape()
    t.cells = flat.ctypes.data_as(C.POINTER(C.c_int32))
    t.cell_off = off.ctypes.data_as(C.POINTER(C.c_int64))
    t.reals = None
    t.real_off = None
    keep = [flat, off]
    if reals_per_env is not None:
        rflat, roff = _flatten(reals_per_env, np.float64)
        t.reals = rflat.ctypes.data_as(C.POINTER(C.c_double))
        t.real_off = roff.ctypes.data_as(C.POINTER(C.c_int64))
        keep += [rflat, roff]
    return t, keep


class Oracle:
    """Batched oracle with the same row model as predpreygrass_b200.BatchedPredPreyGrass."""

    def __init__(self, cfg: PpgConfig, n_envs: int, threads: int = 1):
        self.cfg, self.n_envs = cfg, n_envs
        self.h = lib().ppgo_create(C.byref(cfg), n_envs)
        if not self.h:
            raise ValueError("ppgo_create rejected the config")
        lib().ppgo_set_threads(self.h, threads)
        self.C = (cfg.num_obs_channels + (1 if cfg.variant == 1 and cfg.include_speed_in_obs else 0)
                  + (1 if cfg.variant == 2 and cfg.include_visibility_channel else 0))
        self.grid_C = cfg.num_obs_channels
        self.R = (cfg.obs_range[0], cfg.obs_range[1])
        self._keep = None

    def close(self):
        if self.h:
            lib().ppgo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_tape(self, cells_per_env, reals_per_env=None):
        t, keep = make_tape(cells_per_env, reals_per_env)
        lib().ppgo_load_tape(self.h, C.byref(t))


    def env_reset_eco(self, env, cells, founder_speed):
        c = np.ascontiguousarray(cells, np.int32)
        sp = np.ascontiguousarray(founder_speed if len(founder_speed) else [0.0], np.float64)
        assert lib().ppgo_env_reset_eco(self.h, env, c.ctypes.data, sp.ctypes.data) == 0
        return self.outputs()

    def env_reset_trait(self, env, n_pred, n_prey, cells, founder_trait):
        c = np.ascontiguousarray(cells, np.int32)
        sp = np.ascontiguousarray(founder_trait if len(founder_trait) else [0.0], np.float64)
        assert lib().ppgo_env_reset_trait(self.h, env, int(n_pred), int(n_prey), c.ctypes.data, sp.ctypes.data) == 0
        return self.outputs()

    def read_env_acc(self, env):
        """CAD agent_move_accumulator of one env, in the list order of read_env"""
        st = self.read_env(env)
        acc = [np.zeros(max(1, len(st["ids"][s])), np.float64) for s in range(2)]
        assert lib().ppgo_read_env_acc(self.h, env, acc[0].ctypes.data, acc[1].ctypes.data) == 0
        return tuple(a[: len(st["ids"][s])] for s, a in enumerate(acc))

    def read_env_eco(self, env):
        st = self.read_env(env)
        n = (len(st["ids"][0]), len(st["ids"][1]))
        age = [np.zeros(max(1, n[s]), np.int32) for s in range(2)]
        sp = [np.zeros(max(1, n[s]), np.float64) for s in range(2)]
        dead = np.zeros(max(1, n[1]), np.uint8)
        act = np.zeros(2, np.int32)
        assert lib().ppgo_read_env_eco(self.h, env, age[0].ctypes.data, sp[0].ctypes.data, age[1].ctypes.data,
                                       sp[1].ctypes.data, dead.ctypes.data, act.ctypes.data) == 0
        st.update(age=(age[0][: n[0]], age[1][: n[1]]), speed=(sp[0][: n[0]], sp[1][: n[1]]), dead_prey=dead[: n[1]],
                  active_num=act)
        return st

    def read_episode_eco(self, env):
        """per-episode totals of one ECO env (the dict of BatchedPredPreyGrass.read_episode_eco)"""
        sums, sp = np.zeros(4, np.float64), np.zeros(2, np.int32)
        assert lib().ppgo_read_episode_eco(self.h, env, sums.ctypes.data, sp.ctypes.data) == 0
        return {"distance": (float(sums[0]), float(sums[1])), "move_energy": (float(sums[2]), float(sums[3])), "spawned": (int(sp[0]), int(sp[1]))}

    def read_episode_events_eco(self, env):
        """event counters of one trait-variant env (the dict of BatchedPredPreyGrass.read_episode_events_eco)"""
        ev = np.zeros(6, np.float64)
        assert lib().ppgo_read_episode_events_eco(self.h, env, ev.ctypes.data) == 0
        return {"blocked_capacity": (int(ev[0]), int(ev[1])), "blocked_density": int(ev[2]), "satiation_blocked": int(ev[3]),
                "donated": (float(ev[4]), float(ev[5]))}

    def env_reset_stag(self, env, cells, facing, trait_raw):
        c = np.ascontiguousarray(cells, np.int32)
        f = np.ascontiguousarray(facing if len(facing) else [0], np.int32)
        t = np.ascontiguousarray(trait_raw if len(trait_raw) else [0.0], np.float64)
        assert lib().ppgo_env_reset_stag(self.h, env, c.ctypes.data, f.ctypes.data, t.ctypes.data) == 0
        return self.outputs()

    def read_env_stag(self, env):
        st = self.read_env(env)
        n = (len(st["ids"][0]), len(st["ids"][1]))
        age = [np.zeros(max(1, n[s]), np.int32) for s in range(2)]
        face = np.zeros(max(1, n[0]), np.int32)
        trait = np.zeros(max(1, n[0]), np.float64)
        cap = np.zeros(12, np.int64)
        capr = np.zeros(3, np.float64)
        assert lib().ppgo_read_env_stag(self.h, env, age[0].ctypes.data, face.ctypes.data, trait.ctypes.data, age[1].ctypes.data,
                                        cap.ctypes.data, capr.ctypes.data) == 0
        st.update(age=(age[0][: n[0]], age[1][: n[1]]), facing=face[: n[0]], trait=trait[: n[0]], capture=cap, capture_real=capr)
        return st

    def reset(self, seeds=None, mask=None):
        s = None if seeds is None else np.ascontiguousarray(seeds, np.uint64)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        rc = lib().ppgo_reset(self.h, None if s is None else s.ctypes.data, None if m is None else m.ctypes.data)
        assert rc == 0
        return self.outputs() if mask is None else None

    def step(self, actions_pred, actions_prey):
        a0 = np.ascontiguousarray(actions_pred, np.int32)
        a1 = np.ascontiguousarray(actions_prey, np.int32)
        rc = lib().ppgo_step(self.h, a0.ctypes.data, a1.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"ppgo_step rc={rc}")
        return self.outputs()

    def random_actions(self, seed):
        o = self.outputs()
        a0 = np.zeros(max(1, o["n"][0]), np.int32)
        a1 = np.zeros(max(1, o["n"][1]), np.int32)
        lib().ppgo_random_actions(self.h, seed, a0.ctypes.data, a1.ctypes.data)
        return a0[: o["n"][0]], a1[: o["n"][1]]

    def outputs(self):
        """dict of numpy views (valid until the next call) in the layout of include/ppg.h"""
        b = OracleBuffers()
        lib().ppgo_get_buffers(self.h, C.byref(b))
        n_rows = _np(b.f.n_rows, (4,), np.int32)
        out = {"n_old": (int(n_rows[0]), int(n_rows[1])), "n_new": (int(n_rows[2]), int(n_rows[3]))}
        out["n"] = (out["n_old"][0] + out["n_new"][0], out["n_old"][1] + out["n_new"][1])
        for s in range(2):
            n, R = out["n"][s], self.R[s]
            out[f"obs{s}"] = _np(b.f.obs[s], (n, self.C, R, R), np.float32)
            out[f"obs64_{s}"] = _np(b.obs64[s], (n, self.C, R, R), np.float64)
            out[f"row_env{s}"] = _np(b.f.row_env[s], (n,), np.int32)
            out[f"row_agent{s}"] = _np(b.f.row_agent[s], (n,), np.int32)
            out[f"reward{s}"] = _np(b.f.reward[s], (n,), np.float32)
            out[f"reward64_{s}"] = _np(b.reward64[s], (n,), np.float64)
            out[f"flags{s}"] = _np(b.f.flags[s], (n,), np.uint8)
            out[f"old_off{s}"] = _np(b.f.old_off[s], (self.n_envs + 1,), np.int32)
            out[f"new_off{s}"] = _np(b.f.new_off[s], (self.n_envs,), np.int32)
            out[f"new_cnt{s}"] = _np(b.f.new_cnt[s], (self.n_envs,), np.int32)
        out["env_flags"] = _np(b.f.env_flags, (self.n_envs,), np.uint8)
        out["env_status"] = _np(b.f.env_status, (self.n_envs,), np.uint8)
        out["env_step"] = _np(b.f.env_step, (self.n_envs,), np.int32)
        out["env_count"] = _np(b.f.env_count, (self.n_envs, 2), np.int32)
        return out

    def stats(self):
        out = np.zeros(N_STATS, np.int64)
        lib().ppgo_stats(self.h, out.ctypes.data)
        return out

    def read_env(self, env):
        cap = (max(self.cfg.n_possible[0], 1), max(self.cfg.n_possible[1], 1))
        n = np.zeros(2, np.int32)
        ids = [np.zeros(cap[s], np.int32) for s in range(2)]
        xy = [np.zeros((cap[s], 2), np.int32) for s in range(2)]
        en = [np.zeros(cap[s], np.float64) for s in range(2)]
        gxy = np.zeros((self.cfg.n_grass, 2), np.int32)
        ge = np.zeros(self.cfg.n_grass, np.float64)
        lib().ppgo_read_env(self.h, env, n.ctypes.data, ids[0].ctypes.data, xy[0].ctypes.data, en[0].ctypes.data,
                            ids[1].ctypes.data, xy[1].ctypes.data, en[1].ctypes.data, gxy.ctypes.data, ge.ctypes.data)
        return {
            "ids": (ids[0][: n[0]], ids[1][: n[1]]), "xy": (xy[0][: n[0]], xy[1][: n[1]]),
            "energy": (en[0][: n[0]], en[1][: n[1]]), "grass_xy": gxy, "grass_energy": ge,
        }

    def read_grid(self, env):
        G = self.cfg.grid_size
        g = np.zeros((getattr(self, "grid_C", self.C), G, G), np.float64)
        lib().ppgo_read_grid(self.h, env, g.ctypes.data)
        return g

    # ---- literal single-env interface ----
    def env_reset_cells(self, env, cells):
        c = np.ascontiguousarray(cells, np.int32)
        assert lib().ppgo_env_reset_cells(self.h, env, c.ctypes.data) == 0
        return self.outputs()

    def env_step_ordered(self, env, species, ids, actions):
        s = np.ascontiguousarray(species, np.int32)
        i = np.ascontiguousarray(ids, np.int32)
        a = np.ascontiguousarray(actions, np.int32)
        rc = lib().ppgo_env_step_ordered(self.h, env, len(s), s.ctypes.data, i.ctypes.data, a.ctypes.data)
        if rc != 0:
            raise KeyError("action for an agent that is not alive")
        return self.outputs()

    def env_agents(self, env):
        cap = self.cfg.n_possible[0] + self.cfg.n_possible[1]
        s = np.zeros(cap, np.int32)
        i = np.zeros(cap, np.int32)
        n = lib().ppgo_env_agents(self.h, env, cap, s.ctypes.data, i.ctypes.data)
        return s[:n].copy(), i[:n].copy()
