"""PipelinedPredPreyGrass — the env instances of one GPU split into GROUPS that are stepped on their own CUDA streams.

Why: a step is two kernels of opposite character — the order-dependent state update (`ppg_step_*_kernel`: latency
bound, < 30 % of the issue slots, almost no DRAM traffic) and the observation writer (`ppg_obs_kernel`: a write stream
at 0.6+ of the HBM peak).  Inside ONE handle they are serial per env (rows need the state update), so the device
alternates between an idle memory system and idle schedulers.  Env instances are independent, so two
groups stepped alternately on two streams keep both kinds of work resident at the same time: while group A's rows
stream out, group B's state update runs under them.  This is the double-buffered env layout an RL sampler uses anyway
(policy inference on one half while the other half steps; SURVEY §8b "double-buffer option for overlap with policy
inference").  Trajectories do not depend on the grouping: every env's Philox streams are keyed by its GLOBAL index
(`env_index_base`), exactly as in the multi-GPU sharding (tests/test_gpu_fullsize.py).

The class is plumbing (streams, events, tensor views) over `BatchedPredPreyGrass`; all compute is in the CUDA kernels.
"""
import time

import torch

from . import _lib
from .batched import BatchedPredPreyGrass
from .config import N_STATS, STAT_NAMES


class PipelinedPredPreyGrass:
    def __init__(self, cfg_factory, n_envs, groups=2, device=0):
        """cfg_factory(env_index_base_offset) -> PpgConfig of a group whose env 0 has that offset inside this object."""
        if n_envs % groups:
            raise ValueError("n_envs must be a multiple of groups")
        self.n_envs, self.groups = int(n_envs), int(groups)
        self.per_group = self.n_envs // self.groups
        self.device = torch.device("cuda", device if isinstance(device, int) else device.index or 0)
        self.envs = [BatchedPredPreyGrass(cfg_factory(g * self.per_group), self.per_group, device=self.device.index)
                     for g in range(self.groups)]
        self.cfg = self.envs[0].cfg
        self.C, self.R = self.envs[0].C, self.envs[0].R
        # one stream per group; a single group stays on the caller's current stream
        self.streams = [torch.cuda.Stream(self.device) for _ in range(self.groups)] if self.groups > 1 else [None]
        if self.groups > 1:
            self.envs[0].L.ppg_set_pdl_chain(0)  # parked step kernels would hold the slots the other groups' kernels are meant to fill
        self._ev = [torch.cuda.Event() for _ in range(self.groups)]

    # ------------------------------------------------------------------ plumbing
    def _on(self, g):
        s = self.streams[g]
        return torch.cuda.stream(s) if s is not None else torch.cuda.stream(torch.cuda.current_stream(self.device))

    def fork(self, event=None):
        """record `event` on the current stream and make every group stream wait for it"""
        cur = torch.cuda.current_stream(self.device)
        ev = event if event is not None else torch.cuda.Event()
        ev.record(cur)
        for s in self.streams:
            if s is not None:
                s.wait_event(ev)
        return ev

    def join(self, event=None):
        """the current stream waits for everything queued on the group streams; then records `event`"""
        cur = torch.cuda.current_stream(self.device)
        for g, s in enumerate(self.streams):
            if s is not None:
                self._ev[g].record(s)
                cur.wait_event(self._ev[g])
        if event is not None:
            event.record(cur)
        return event

    def close(self):
        for e in self.envs:
            e.close()
        self.envs = []

    # ------------------------------------------------------------------ env API, per group
    def reset(self):
        self.fork()
        for g, e in enumerate(self.envs):
            with self._on(g):
                e.reset()
        self.join()
        return [e.out for e in self.envs]

    def step_group(self, g, actions_pred=None, actions_prey=None):
        """step group g on its stream (actions: int32 CUDA tensors indexed by the rows of the group's last output)"""
        with self._on(g):
            return self.envs[g].step(actions_pred, actions_prey)

    def rollout_random(self, k, seed):
        """k lockstep steps of every group with uniform random actions from the device-side Philox generator
        (`ppg_rollout_random`: the launches of the groups are issued alternately from C, each group on its stream, so
        that they overlap on the device and the Python interpreter is not in the launch path)"""
        import ctypes as C

        n = self.groups
        hs = (C.c_void_p * n)(*[e.h for e in self.envs])
        cur = torch.cuda.current_stream(self.device).cuda_stream
        st = (C.c_void_p * n)(*[(s.cuda_stream if s is not None else cur) for s in self.streams])
        _lib.check(self.envs[0].L.ppg_rollout_random(hs, n, st, int(k), int(seed)), self.envs[0].h)

    # ------------------------------------------------------------------ statistics
    def stats_device(self):
        self.fork()
        parts = []
        for g, e in enumerate(self.envs):
            with self._on(g):
                parts.append(e.stats_device().clone())
        self.join()
        tot = parts[0]
        for p in parts[1:]:
            tot = tot + p
        return tot

    def stats(self):
        v = self.stats_device().tolist()
        for e in self.envs:
            e.stats()  # raises if a device error word is set
        d = dict(zip(STAT_NAMES, v))
        assert len(v) == N_STATS
        return d

    def launch_count(self):
        return sum(e.launch_count() for e in self.envs)

    # ------------------------------------------------------------------ measurement helpers (bench.py)
    def profile_rollout(self, k, seed):
        """per-kernel CUDA-event timing: k steps of every group, one group after the other, kernels back to back.
        -> (ms in the step kernels, ms in the observation kernels), summed over all launches"""
        a = b = 0.0
        for e in self.envs:
            torch.cuda.synchronize(self.device)
            e.profile_begin()
            for _ in range(k):
                a0, a1 = e.random_actions(seed)
                e.step(a0, a1)
            x, y, _ = e.profile_end()
            a += x
            b += y
        torch.cuda.synchronize(self.device)
        return a, b

    def bench_host(self, steps, pools, row_bytes, barrier, world, dev):
        """e2e through `ppg_step_host`: actions from pinned host memory, the whole row batch back into pinned host memory,
        every step, for every group (the groups' calls are issued one after the other on their own streams)."""
        import torch.distributed as dist

        hosts, pool, n = [], [], []
        for e in self.envs:
            hosts.append(e.make_host_buffers(pinned=True))
            p0, p1 = pools(max(e.row_capacity) + 4096)
            pool.append((torch.from_numpy(p0).pin_memory(), torch.from_numpy(p1).pin_memory()))
            n.append(list(e.out.counts()))
        h2d = d2h = 0
        e0 = self.stats_device()
        barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            off = (i * 61) % 4096
            for g, e in enumerate(self.envs):
                n0, n1 = n[g]
                hosts[g]["actions0"][:n0].copy_(pool[g][0][off:off + n0])  # host -> pinned staging of this step's inputs
                hosts[g]["actions1"][:n1].copy_(pool[g][1][off:off + n1])
                h2d += 4 * (n0 + n1)
                with self._on(g):
                    n0, n1 = e.step_host(hosts[g])
                n[g] = [n0, n1]
                d2h += n0 * (row_bytes[0] + 13) + n1 * (row_bytes[1] + 13) + 4 * (e.n_envs + 1) * 4 + e.n_envs * 14 + 16
        torch.cuda.synchronize(self.device)
        dt = time.perf_counter() - t0
        ed = (self.stats_device() - e0).to(torch.float64)
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ed, op=dist.ReduceOp.SUM)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        gbs = d2h / dt / 1e9
        return {"value": float(ed[1].item()) / float(tt.item()), "h2d_bytes_per_step": h2d // steps, "d2h_bytes_per_step": d2h // steps,
                "steps": steps, "d2h_gbs_this_rank": gbs,
                "api": "ppg_step_host (pinned host actions in, full row batch incl. observations out)"}

    def bench_device_policy(self, steps, barrier, world, dev):
        """e2e on the zero-copy consumer path (`connector.DeviceRollout`): a torch policy reads the row batch in place,
        writes the actions in place, the env steps — no host copy of observations, one 16-byte count read per group and step."""
        import torch.distributed as dist

        from .connector import DeviceRollout, LinearPolicy

        runners = []
        for g, e in enumerate(self.envs):
            with self._on(g):
                pol = [LinearPolicy(self.C * self.R[s] ** 2, e.n_actions(s), device=self.device, seed=7 + s) for s in range(2)]
                runners.append(DeviceRollout(e, pol, stream=self.streams[g]))
        for r in runners:
            r.step()
        e0 = self.stats_device()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            for r in runners:
                r.step()
        torch.cuda.synchronize(self.device)
        dt = time.perf_counter() - t0
        ed = (self.stats_device() - e0).to(torch.float64)
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ed, op=dist.ReduceOp.SUM)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return {"value": float(ed[1].item()) / float(tt.item()), "steps": steps, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 16 * self.groups,
                "api": "connector.DeviceRollout: torch policy (one linear layer per species, sampled actions) on the row batch in place, "
                       "actions written in place, ppg_step; host wall clock"}
