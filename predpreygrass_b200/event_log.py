"""Host-side exporters of eco_evolutionary: `per_step_agent_data` (ECO:426-446) and `agent_event_log` (ECO:1488-1500) rebuilt
from what the device hands back every step.

The reference fills both inside `step()` while it walks its dicts.  The device step keeps none of this (it is analytics of
the evaluation scripts and the renderer's tooltips), but everything in it follows from the state before the step,
the state after it and the row flags, because the energy bookkeeping of ECO is a fixed chain per agent:

    E0 --(-basal loss, ECO:597-599)--> E1 --(-move cost, ECO:644-647)--> E2 --(+bite, ECO:812-818 / 903-908)--> E3
       --(-offspring energy, ECO:1147-1151 / 1243-1247)--> E4

Each link is the reference's own float64 expression evaluated on the host with the same operands, so the deltas come out
bit-identical; the chain's end is compared with the energy the device reports and a mismatch is counted in
`inexact_chains` (0 on every recording of tests/golden/eco_events_*.json.gz).  The only thing the device does not report is
where an agent that died this step stood when it died; its move is therefore replayed on the host (target cell, or the old
cell if the own-species layer blocked it), a fully eaten prey is matched to its predator through that cell, and the match is
confirmed by the predator's energy.

The per-agent records of the reference (`agent_stats_live` / `agent_stats_completed`, read by the evaluation scripts through
`get_all_agent_stats()`, ECO:1676-1682) are rebuilt alongside, field by field — distance, locomotion energy, meals, average energy,
cumulative reward (with the step's reward that `_finalize_agent_record` adds once more, ECO:1540-1542), frozen death steps.
Whether the last move of an agent that dies in a step was blocked is replayed on the own-species layer the reference tests
(ECO:684-686).  Lineage `reward_events` are logged with the configured coefficient.
"""
import json

import numpy as np

ROW_TERMINATED, ROW_TRUNCATED, ROW_ATE, ROW_CARCASS, ROW_REPRODUCED = 0x01, 0x02, 0x10, 0x20, 0x40


def _role(value, agent):
    """`_get_role_specific` (ECO:1723-1730): scalars, or dicts keyed by an agent-id prefix"""
    if isinstance(value, dict):
        for k in value:
            if agent.startswith(k):
                return value[k]
        raise KeyError(f"Type-specific key '{agent}' not found")
    return value


class EcoEventRecorder:
    def __init__(self, config, action_to_move, grid_size, speed_distance_threshold):
        g = config.get
        self.cfg = config
        self.loss = (g("energy_loss_per_step_predator"), g("energy_loss_per_step_prey"))
        self.move_cost = (float(g("movement_energy_cost_per_cell_predator", 0.0)), float(g("movement_energy_cost_per_cell_prey", 0.0)))
        self.exponent = float(g("movement_speed_cost_exponent", 2.0))  # ECO:81
        self.grass_gain, self.grass_max = g("energy_gain_per_step_grass"), g("max_energy_grass")
        self.cap_prey = float(g("max_energy_gain_per_prey", float("inf")))
        self.cap_grass = float(g("max_energy_gain_per_grass", float("inf")))
        self.init_e = (float(g("initial_energy_predator")), float(g("initial_energy_prey")))
        self.max_age = {"predator": 120, "prey": 100}  # ECO:52-63: defaults for the roles a configured dict leaves out
        if isinstance(g("max_agent_age"), dict):
            self.max_age.update(g("max_agent_age"))
        self.carcass_age = g("carcass_only_predator_age", None)
        self.moves, self.G, self.thr = action_to_move, int(grid_size), float(speed_distance_threshold)
        self.jump_low, self.jump_high = int(g("slow_max_move_distance", 1)), int(g("fast_max_move_distance", 2))  # ECO:551-557
        self.reset({}, [], {})

    # ---------------------------------------------------------------- helpers (each one the reference's own expression)
    def _limit(self, caps, agent):
        if caps is None:
            return None
        if isinstance(caps, dict):
            for prefix, limit in caps.items():
                if agent.startswith(prefix):
                    return limit
            return None
        return caps

    def _cost(self, agent, s, old, new, speed):
        """`_get_movement_energy_cost` (ECO:565-573)"""
        distance = float(np.linalg.norm(np.array(new) - np.array(old)))
        if distance <= 0:
            return 0.0
        factor = 1.0 if speed is None else float(speed) ** self.exponent  # ECO:559-563
        return self.move_cost[s] * distance * factor

    def _target(self, pos, action, speed):
        """`_get_move` without the occupancy test (ECO:664-686): the cell the agent moves to unless it is blocked"""
        mv = self.moves[int(action)]
        max_d = self.jump_low if (speed is None or float(speed) < self.thr) else self.jump_high
        if max(abs(mv[0]), abs(mv[1])) > max_d:
            mv = (int(np.sign(mv[0])) * max_d, int(np.sign(mv[1])) * max_d)
        return (min(max(pos[0] + mv[0], 0), self.G - 1), min(max(pos[1] + mv[1], 0), self.G - 1))

    # ---------------------------------------------------------------- episode start
    def reset(self, state, agents, grass):
        """state: {agent: (pos, energy, age, speed or None, dead)}; grass: {cell: (name, energy)}"""
        self.per_step_agent_data = []
        self.agent_event_log = {}
        self.agent_parents, self.agent_offspring_counts, self.agent_live_offspring_ids = {}, {}, {}
        self.cumulative_reward = {}
        self.lineage = {}  # agent -> [parent, live descendants, live descendants at the last reward pass, counted as alive]
        self.first_bite_step = {}
        self.inexact_chains = 0
        self.closed = set()  # agents whose record is finalized (no longer in `agent_stats_live`)
        self.stats, self.step_reward = {}, {}  # `agent_stats_live` + `agent_stats_completed` records; `self.rewards` of the running step
        self.live_order, self.completed_order = [], []
        self.ambiguous_final_moves = 0  # agents that starved on a step whose move may or may not have been blocked (see get_all_agent_stats)
        self.prev, self.prev_agents, self.prev_grass = dict(state), list(agents), dict(grass)
        for a in agents:
            self._register(a, None, 0, state[a][3])

    def _register(self, agent, parent, t, speed):
        self.agent_parents[agent] = parent
        self.agent_offspring_counts[agent] = 0
        self.agent_live_offspring_ids[agent] = []
        self.cumulative_reward[agent] = 0.0
        self.lineage[agent] = [parent, 0, 0, False]
        role = "predator" if "predator" in agent else "prey"
        self.live_order.append(agent)
        self.stats[agent] = {  # `agent_stats_live[agent]` (ECO:1501-1523)
            "agent_id": agent, "birth_step": t, "parent": parent, "offspring_count": 0, "offspring_ids": self.agent_live_offspring_ids[agent],
            "distance_traveled": 0.0, "movement_energy_spent": 0.0, "times_ate": 0, "energy_gained": 0.0, "avg_energy_sum": 0.0,
            "avg_energy_steps": 0, "cumulative_reward": 0.0, "lineage_reward_total": 0.0, "policy_group": role,
            "genome": None if speed is None else {"speed": float(speed)}, "death_step": None, "death_cause": None, "avg_energy": 0.0,
            "max_age": self._limit(self.max_age, agent), "age_expired_step": None, "carcass_only_blocks": 0}
        self._lineage_alive(agent, True)  # `_handle_lineage_birth` (ECO:1462-1466)
        self.agent_event_log[agent] = {
            "agent_id": agent, "birth_step": t, "death_step": None, "parent_id": parent, "death_cause": None,
            "eating_events": [], "reproduction_events": [], "reward_events": [], "diet_events": [], "lifecycle_events": [],
            "genome": None if speed is None else {"speed": float(speed)},
        }

    def _lineage_alive(self, agent, alive):
        """`_set_lineage_alive_flag` + `_propagate_lineage_delta` (ECO:1440-1460): every ancestor, dead or alive, counts the
        agent among its live descendants while the flag is set"""
        rec = self.lineage[agent]
        if rec[3] == alive:
            return
        rec[3] = alive
        cur = rec[0]
        while cur is not None and cur in self.lineage:
            self.lineage[cur][1] += 1 if alive else -1
            cur = self.lineage[cur][0]

    def _finalize(self, agent, cause, step):
        """the event-log part of `_finalize_agent_record` (ECO:1532-1573); every death path of the reference calls
        `_handle_lineage_death` first (ECO:764, 1072; a bitten prey at its first bite, ECO:840, 846)"""
        self._lineage_alive(agent, False)
        self.closed.add(agent)
        if agent in self.live_order:  # the record part of `_finalize_agent_record` (ECO:1536-1551)
            self.live_order.remove(agent)
            self.completed_order.append(agent)
            rec = self.stats[agent]
            rec["cumulative_reward"] += self.step_reward.get(agent, 0.0)  # the step's reward is added once more (ECO:1540-1542)
            rec["death_cause"] = cause
            if rec["death_step"] is None:
                rec["death_step"] = step
            rec["offspring_count"] = self.agent_offspring_counts.get(agent, rec["offspring_count"])
            rec["avg_energy"] = rec["avg_energy_sum"] / max(rec["avg_energy_steps"], 1)
        evt = self.agent_event_log[agent]
        evt["death_step"] = self.first_bite_step.get(agent, step)
        evt["death_cause"] = cause
        self.agent_live_offspring_ids.pop(agent, None)

    # ---------------------------------------------------------------- one step
    def step(self, t, action_dict, rows, state, newborn, grass, time_limit):
        """t: `current_step` during the step; rows: {agent: flags} of every row of the step; state / grass: after the step;
        newborn: this step's newborns in birth order, predators first (the tail of the reference's `self.agents`)"""
        cfg, prev = self.cfg, self.prev
        agents = [a for a in self.prev_agents if a in state] + list(newborn)  # ECO:353-367
        is_pred = lambda a: "predator" in a  # noqa: E731
        E, deltas, age = {}, {}, {}
        self.step_reward = {}
        moved = {}  # dying agents: per candidate cell (distance, cost, energy after the move) for the record's movement totals
        gone = set()  # terminated so far in this step
        # Step 1 (ECO:582-614): basal loss, ageing, age cap
        for a in self.prev_agents:
            s = 0 if is_pred(a) else 1
            pos, e0, ag, spd, dead = prev[a]
            E[a] = e0 - self.loss[s]
            deltas[a] = {"decay": -self.loss[s], "move": 0.0, "eat": 0.0, "repro": 0.0}
            age[a] = ag
            if not (s == 1 and dead):
                age[a] = ag + 1
        for a in self.prev_agents:
            s = 0 if is_pred(a) else 1
            limit = self._limit(self.max_age, a)
            if not (s == 1 and prev[a][4]) and isinstance(limit, (int, float)) and limit >= 0 and age[a] >= limit:
                self.agent_event_log[a]["lifecycle_events"].append({"t": int(t), "event": "max_age_reached", "age": int(age[a])})
                self.stats[a]["age_expired_step"] = int(t)  # ECO:1076-1078
                self.step_reward[a] = 0.0
                self._finalize(a, "max_age", t)
                gone.add(a)
        # Step 2 (ECO:616-624): grass regrowth
        g_now = {cell: (name, min(e + self.grass_gain, self.grass_max)) for cell, (name, e) in self.prev_grass.items()}
        # Step 3 (ECO:626-662): moves.  Survivors: old and new cell are known.  Agents that die in this step: the cells they
        # can have died on are kept as candidates (cell, energy) — the move target, or the old cell if the move was blocked;
        # terminated agents (age cap) and carcasses do not move.
        # Whether the move of an agent that dies in this step was blocked is replayed on the own-species layer the reference tests
        # (`grid_world_state[layer, target] > 0`, float32, ECO:684-686): every agent writes its cell after the basal loss
        # (carcasses included, aged-out agents zero theirs), movers zero their old cell and write the new one.
        cell = ({}, {})
        for a in self.prev_agents:
            cell[0 if is_pred(a) else 1][prev[a][0]] = np.float32(0.0 if a in gone else E[a])
        cand = {}
        for a, action in action_dict.items():
            if a not in prev or a in gone or (not is_pred(a) and prev[a][4]):
                continue
            s = 0 if is_pred(a) else 1
            old, spd = prev[a][0], prev[a][3]
            if a in state:
                new = state[a][0]
                cost = self._cost(a, s, old, new, spd)
                E[a] -= cost
                deltas[a]["move"] -= cost
                self._moved(a, old, new, cost, E[a])
                e_after = E[a]
            else:
                new = self._target(old, action, spd)
                if cell[s].get(new, 0.0) > 0:
                    new = old
                c_t = self._cost(a, s, old, new, spd)
                e_after = E[a] - c_t
                cand[a] = [(new, e_after)]
                moved[a] = {new: (old, new, c_t, e_after)}
            cell[s][old] = np.float32(0.0)
            cell[s][new] = np.float32(e_after)
        # Step 4a (ECO:314-319): starvation.  Predators that are gone and not aged out starved.  A prey that is gone stays
        # on the grid until Step 5, so a predator can still bite it (ECO:789-791 looks at `agent_positions`), whatever it
        # died of: every prey that is gone keeps its candidates for Step 4c.
        dead_prey = {}
        for a in self.prev_agents:
            if a in state:
                continue
            if a not in cand:  # aged out in Step 1 or a carcass: it did not move
                cand[a] = [(prev[a][0], E[a])]
            if is_pred(a):
                if a not in gone:
                    if a in moved:  # which of its candidate cells it starved on: the one that leaves no energy
                        opts = [v for v in moved[a].values() if v[3] <= 0] or list(moved[a].values())
                        if len(opts) > 1:
                            self.ambiguous_final_moves += 1
                        self._moved(a, *opts[0])
                    self.step_reward[a] = 0  # ECO:768
                    self._finalize(a, "starved", t)
                    gone.add(a)
            else:
                dead_prey[a] = [{"cell": c, "e": e, "grass": None} for c, e in cand[a]]
        # Step 4b (ECO:886-939): prey eat grass (not the carcasses, not the terminated ones)
        for a in self.prev_agents:
            if is_pred(a) or a in gone:
                continue
            if prev[a][4]:
                if a in state or E[a] > 0:  # a carcass that starved was terminated in Step 4a, before the prey engagements
                    self._reward(a, _role(cfg.get("reward_prey_step", 0.0), a))
                continue
            f = rows.get(a, 0)
            if a in state and f & ROW_ATE and state[a][0] in g_now:
                name, ge = g_now[state[a][0]]
                bite = min(float(ge), self.cap_grass)
                E[a] += bite
                deltas[a]["eat"] = bite
                self._reward(a, _role(cfg.get("reward_prey_eat_grass", 0.0), a))
                self.stats[a]["times_ate"] += 1
                self.stats[a]["energy_gained"] += bite
                self.agent_event_log[a]["eating_events"].append({"t": int(t), "id_eaten": name, "alive_before_bite": True,
                                                                 "bite_size": float(bite), "energy_after": float(E[a])})
            elif a in dead_prey and f & ROW_ATE:
                # ate grass, then was eaten in the same step: on a candidate cell with a patch (and energy left: a starved
                # prey is terminated before it can eat) the bite raises that candidate's energy
                for c in dead_prey[a]:
                    if c["e"] > 0 and c["cell"] in g_now:
                        name, ge = g_now[c["cell"]]
                        bite = min(float(ge), self.cap_grass)
                        c["e"] += bite
                        c["grass"] = (name, bite)
            elif a in state:
                self._reward(a, _role(cfg.get("reward_prey_step", 0.0), a))
        # Step 4c (ECO:775-884): predators
        prey_at = {state[a][0]: a for a in state if not is_pred(a) and a in prev}
        eaten = {}
        for a in self.prev_agents:
            if not is_pred(a) or a in gone or a not in state:
                continue
            f = rows.get(a, 0)
            pos = state[a][0]
            gone_here = [(q, c) for q, cs in dead_prey.items() if q not in eaten for c in cs if c["cell"] == pos]
            if not f & ROW_ATE:
                q = prey_at.get(pos)
                was_dead = prev[q][4] if q is not None else False
                if q is None and gone_here:
                    q, c = gone_here[0]
                    was_dead = prev[q][4] and c["e"] > 0  # a carcass that starved left `dead_prey` when it did (ECO:761-763)
                limit = self._limit(self.carcass_age, a)
                if q is not None and not was_dead and isinstance(limit, (int, float)) and limit >= 0 and age[a] < limit:
                    self.agent_event_log[a]["diet_events"].append({"t": int(t), "event": "carcass_only_block", "prey_id": q, "age": int(age[a])})  # ECO:1035-1059
                    self.stats[a]["carcass_only_blocks"] += 1
                self._reward(a, _role(cfg.get("reward_predator_step", 0.0), a))
                continue
            # which prey: a bitten one that is still there (carcass), else a prey that is gone and can have stood here
            q = prey_at.get(pos)
            if q is not None and state[q][4]:
                was_dead = prev[q][4]
                pe = E[q]
                bite = min(float(pe), self.cap_prey)
                E[q] = pe - bite
                if q not in self.first_bite_step:
                    self.first_bite_step[q] = int(t)  # ECO:836-838: the first bite freezes death_step
                if self.stats[q]["death_step"] is None:
                    self.stats[q]["death_step"] = int(t)
                self._lineage_alive(q, False)  # ECO:839-840
            else:
                if not gone_here:
                    self.inexact_chains += 1
                    continue
                want = state[a][1] + (self.init_e[0] if f & ROW_REPRODUCED else 0.0)  # only used to choose between candidates
                q, c = min(gone_here, key=lambda o: abs((E[a] + min(float(o[1]["e"]), self.cap_prey)) - want))
                was_dead = prev[q][4] and c["e"] > 0  # a carcass that starved left `dead_prey` when it did (ECO:761-763)
                bite = min(float(c["e"]), self.cap_prey)
                eaten[q] = c
            E[a] += bite
            deltas[a]["eat"] = bite
            self._reward(a, _role(cfg.get("reward_predator_catch_prey", 0.0), a))
            self.stats[a]["times_ate"] += 1
            self.stats[a]["energy_gained"] += bite
            self.agent_event_log[a]["eating_events"].append({"t": int(t), "id_eaten": q, "alive_before_bite": not was_dead,
                                                             "bite_size": float(bite), "energy_after": float(E[a])})
        for q, cs in dead_prey.items():
            c = eaten.get(q)
            if c is not None and c["grass"] is not None:  # its last meal (ECO:926-937)
                self.agent_event_log[q]["eating_events"].append({"t": int(t), "id_eaten": c["grass"][0], "alive_before_bite": True,
                                                                 "bite_size": float(c["grass"][1]), "energy_after": float(c["e"])})
            if q in gone:  # aged out: the cause stays (its record was closed in Step 1)
                continue
            was_eaten = c is not None and c["e"] > 0
            if q in moved:  # the record's movement totals: the cell it was caught on, else the one it starved on
                if was_eaten:
                    opts = [v for k, v in moved[q].items() if k == c["cell"]]
                else:
                    opts = [v for v in moved[q].values() if v[3] <= 0] or list(moved[q].values())
                    if len(opts) > 1:
                        self.ambiguous_final_moves += 1
                self._moved(q, *opts[0])
            if was_eaten:
                if not prev[q][4]:  # alive at the prey engagements (Step 4b): its meal or the step reward
                    if c["grass"] is not None:
                        self._reward(q, _role(cfg.get("reward_prey_eat_grass", 0.0), q))
                        self.stats[q]["times_ate"] += 1
                        self.stats[q]["energy_gained"] += c["grass"][1]
                    else:
                        self._reward(q, _role(cfg.get("reward_prey_step", 0.0), q))
                penalty = _role(cfg.get("penalty_prey_caught", 0.0), q)  # ECO:851-861
                self.step_reward[q] = penalty
                self.stats[q]["cumulative_reward"] += penalty
                self._finalize(q, "eaten", t)
            else:
                if c is None and all(x["e"] > 0 for x in cs):
                    self.inexact_chains += 1  # gone, but it can neither have starved nor did a predator take it
                self.step_reward[q] = 0  # ECO:768
                self._finalize(q, "starved", t)
            gone.add(q)
        # Step 6 (ECO:1092-1270): births, predators first; the k-th newborn row of a species belongs to the k-th parent
        new_agents = list(newborn)
        for s, role in enumerate(("predator", "prey")):
            parents = [a for a in self.prev_agents if (is_pred(a) == (s == 0)) and a in state and rows.get(a, 0) & ROW_REPRODUCED]
            children = [a for a in new_agents if is_pred(a) == (s == 0)]
            if len(parents) != len(children):
                self.inexact_chains += 1
            for par, child in zip(parents, children):
                self._register(child, par, int(t), state[child][3])
                self.agent_live_offspring_ids[par].append(child)
                self.agent_offspring_counts[par] += 1
                self.agent_event_log[par]["reproduction_events"].append({"t": int(t), "child_id": child})
                E[par] -= self.init_e[s]
                deltas[par]["repro"] = -self.init_e[s]
                self.stats[par]["offspring_count"] += 1
                self.step_reward[child] = 0
                r = _role(cfg.get(f"reproduction_reward_{role}", 0.0), par)
                self._reward(par, r)
                self.agent_event_log[par]["reward_events"].append({"t": int(t), "reproduction_reward": float(r), "lineage_reward": 0.0,
                                                                   "cumulative_reward": float(self.cumulative_reward[par])})
                deltas[child] = {"decay": 0.0, "move": 0.0, "eat": 0.0, "repro": 0.0}
        # Step 6.5 (ECO:943-984): lineage survival rewards of the agents whose record is still open; an event is logged for
        # every change of the live-descendant count, also when the coefficient (and with it the reward) is zero
        for a, rec in self.lineage.items():
            if a in self.closed:
                continue
            delta = rec[1] - rec[2]
            rec[2] = rec[1]
            if delta == 0:
                continue
            reward = _role(cfg.get("lineage_reward_coeff", 0.0), a) * float(delta)
            self.stats[a]["lineage_reward_total"] += reward  # ECO:962-964
            if reward != 0:
                self.cumulative_reward[a] += reward
                self.stats[a]["cumulative_reward"] += reward
                self.step_reward[a] = self.step_reward.get(a, 0.0) + reward
            self.agent_event_log[a]["reward_events"].append({"t": int(t), "reproduction_reward": 0.0, "lineage_reward": float(reward),
                                                             "cumulative_reward": float(self.cumulative_reward[a])})
        # the chain's end is the device's energy
        for a in self.prev_agents:
            if a in state and E[a] != state[a][1]:
                self.inexact_chains += 1
        # per_step_agent_data (ECO:426-446); `offspring_ids` is the live list object itself, as in the reference
        step_data = {}
        for a in agents:
            pos, e, ag, spd, dead = state[a]
            d = deltas[a]
            step_data[a] = {"position": pos, "energy": e, "energy_decay": d["decay"], "energy_movement": d["move"],
                            "energy_eating": d["eat"], "energy_reproduction": d["repro"], "age": ag,
                            "offspring_count": self.agent_offspring_counts[a],
                            "offspring_ids": self.agent_live_offspring_ids.get(a, []), "parent": self.agent_parents.get(a)}
        self.per_step_agent_data.append(step_data)
        if time_limit:  # ECO:478-479: every record still open is closed with the step counter already advanced
            for a in list(self.live_order):
                self._finalize(a, "time_limit", int(t) + 1)
        self.prev, self.prev_agents, self.prev_grass = dict(state), list(agents), dict(grass)

    def _reward(self, agent, r):
        """`self.rewards[agent] = r` plus the running totals of the record and of the reward events"""
        self.step_reward[agent] = r
        self.cumulative_reward[agent] += r
        if agent in self.stats:
            self.stats[agent]["cumulative_reward"] += r

    def _moved(self, agent, old, new, cost, energy_after):
        """the record's part of a move (ECO:656-660)"""
        rec = self.stats[agent]
        rec["avg_energy_sum"] += energy_after
        rec["avg_energy_steps"] += 1
        rec["distance_traveled"] += float(np.linalg.norm(np.array(new) - np.array(old)))
        rec["movement_energy_spent"] += cost

    def record_order(self):
        return list(self.live_order) + list(self.completed_order)

    def get_all_agent_stats(self):
        """`get_all_agent_stats` (ECO:1676-1682).  The records follow the reference's bookkeeping field by field; the one thing the
        device does not report is whether the last move of an agent that STARVED was blocked — if both of its candidate cells
        leave it without energy the record books the move as made and `ambiguous_final_moves` counts it."""
        out = {}
        for a in self.record_order():
            rec = dict(self.stats[a])
            rec["offspring_ids"] = list(rec["offspring_ids"])
            out[a] = rec
        return out

    def get_total_offspring_by_type(self):
        counts = {"predator": 0, "prey": 0}
        for a in self.record_order():
            counts[self.stats[a]["policy_group"]] += self.stats[a]["offspring_count"]
        return counts

    def export(self, path):
        """`export_agent_event_log` (ECO:1575-1597)"""
        with open(path, "w", encoding="utf-8") as f:
            json.dump(self.agent_event_log, f, indent=2)


class TraitEventRecorder:
    """The same exporters for the trait variants metabolic_rate / investment (MR:391-411, 1153-1165): their step has no
    carcasses, no age caps and no lineage rewards, and a fixed chain per agent

        E0 --(-basal cost [x rate], MR:555-563)--> E1 --(-move cost, MR:516-523)--> E2 --(+gain [x rate ** alpha], MR:747-751,
           805-807)--> E3 --(-offspring energy [parent energy x fraction], MR:897-899 / INV:546-557)--> E4

    evaluated with the reference's own float64 expressions.  Besides the two exporters the recorder keeps the order in which
    the reference iterates its agent records (`_iter_all_agent_records`, MR:1400-1404: live records in registration order,
    then the completed ones in the order they were finalized) — the `*_repro_spearman` metrics rank with argsort and no tie
    correction, so they depend on it (MR:1358-1382)."""

    def __init__(self, config, action_to_move, grid_size, trait):
        g = config.get
        self.cfg, self.trait = config, trait
        mr, inv = trait == "metabolic_rate", trait == "offspring_investment_fraction"
        self.cad = trait == "speed"  # eco_evolutionary_cadence: the speed genome sets the move rate
        if self.cad:
            self.speed_coeff = float(g("metabolic_speed_coeff", 1.0))  # CAD:79
            self.exponent = float(g("movement_speed_cost_exponent", 2.0))
            self.max_cooldown = g("max_cooldown", 10)  # CAD:114
            self.max_age = {"predator": 120, "prey": 100}  # CAD:59-70
            if isinstance(g("max_agent_age"), dict):
                self.max_age.update(g("max_agent_age"))
            self.cap_grass = float(g("max_energy_gain_per_grass", float("inf")))  # CAD:912
            self.record_step_data = bool(g("record_step_data", False))  # CAD:83,422
        if inv or self.cad:
            self.loss = (g("energy_loss_per_step_predator"), g("energy_loss_per_step_prey"))  # INV:54-55
        else:
            self.loss = (g("basal_energy_cost_predator"), g("basal_energy_cost_prey"))  # MR:50-51
        self.scale_loss = mr
        self.alpha = float(g("metabolic_rate_alpha", 0.7)) if mr else None  # MR:102
        self.move_cost = (float(g("movement_energy_cost_per_cell_predator", 0.0)), float(g("movement_energy_cost_per_cell_prey", 0.0)))
        self.grass_gain, self.grass_max = g("energy_gain_per_step_grass"), g("max_energy_grass")
        self.cap_prey = float(g("max_energy_gain_per_prey", float("inf")))  # MR:64
        self.init_e = None if inv else (g("initial_energy_predator"), g("initial_energy_prey"))
        self.coop = trait == "cooperation_rate"
        if self.coop or self.cad:
            self.cap_prey = float("inf")  # COOP:776-777, CAD:856-857: the whole prey
        self.coop_range = int(g("cooperation_range", 2))  # COOP:93
        self.sat_cd = -1 if (self.coop or self.cad) else int(g("predator_satiation_cooldown", 0))  # MR:62
        self.moves, self.G = action_to_move, int(grid_size)
        self.reset({}, [], {})

    def _cost(self, s, old, new, speed=None):
        """`_get_movement_energy_cost` (MR:516-523; CAD:598-606 with the factor speed ** exponent)"""
        distance = float(np.linalg.norm(np.array(new) - np.array(old)))
        if distance <= 0:
            return 0.0
        if self.cad:
            return self.move_cost[s] * distance * (1.0 if speed is None else float(speed) ** self.exponent)
        return self.move_cost[s] * distance

    def _move_rate(self, speed):
        """`_get_agent_move_rate` (CAD:556-578)"""
        if speed is None:
            return 1.0
        normalized = max(0.0, min(1.0, float(speed)))
        min_rate = 1.0 / float(self.max_cooldown)
        return min_rate + normalized * (1.0 - min_rate)

    def _target(self, pos, action):
        """`_get_move` without the occupancy test (MR:617-630)"""
        mv = self.moves[int(action)]
        return (min(max(pos[0] + mv[0], 0), self.G - 1), min(max(pos[1] + mv[1], 0), self.G - 1))

    def _gain(self, food, rate):
        """MR:751 / 807: food x rate ** alpha; INV:759 / 812: the food itself"""
        if self.alpha is None:
            return food
        return food * ((float(rate) if rate is not None else 1.0) ** self.alpha)

    def reset(self, state, agents, grass):
        """state: {agent: (pos, energy, age, trait value or None, _)}; grass: {cell: (name, energy)}"""
        self.per_step_agent_data = []
        self.agent_event_log = {}
        self.agent_parents, self.agent_offspring_counts, self.agent_live_offspring_ids = {}, {}, {}
        self.cumulative_reward = {}
        self.live_order, self.completed_order = [], []  # `agent_stats_live` / `agent_stats_completed` key order
        self.stats = {}        # the records themselves
        self.step_reward = {}  # `self.rewards` of the running step (what `_finalize_agent_record` adds once more, MR:1202-1204)
        self.sat_until = {}  # agent_satiation_until (MR:134)
        self.energy_donated, self.kin_donation = [0.0, 0.0], [0.0, 0.0]  # total_energy_donated / total_kin_donation (COOP:159-162)
        self.inexact_chains = 0
        self.prev, self.prev_agents, self.prev_grass = dict(state), list(agents), dict(grass)
        for a in agents:
            self._register(a, None, 0, state[a][3])

    def _register(self, agent, parent, t, value):
        self.agent_parents[agent] = parent
        self.agent_offspring_counts[agent] = 0
        self.agent_live_offspring_ids[agent] = []
        self.cumulative_reward[agent] = 0.0
        self.live_order.append(agent)
        self.agent_event_log[agent] = {
            "agent_id": agent, "birth_step": t, "death_step": None, "parent_id": parent, "death_cause": None,
            "eating_events": [], "reproduction_events": [], "reward_events": [], "diet_events": [], "lifecycle_events": [],
            "genome": None if value is None else {self.trait: float(value)},
        }
        if self.cad:
            del self.agent_event_log[agent]["diet_events"]  # CAD:1327-1338
        # `agent_stats_live[agent]` (MR:1166-1192, COOP:1183-1185, CAD:1339-1362)
        rec = {"agent_id": agent, "birth_step": t, "parent": parent, "offspring_count": 0, "offspring_ids": self.agent_live_offspring_ids[agent],
               "distance_traveled": 0.0, "movement_energy_spent": 0.0, "times_ate": 0, "energy_gained": 0.0, "avg_energy_sum": 0.0,
               "avg_energy_steps": 0, "cumulative_reward": 0.0, "policy_group": "predator" if "predator" in agent else "prey",
               "genome": None if value is None else {self.trait: float(value)}, "death_step": None, "death_cause": None, "avg_energy": 0.0}
        if self.cad:
            rec.update(max_age=self.max_age.get(rec["policy_group"]), age_expired_step=None)
        else:
            rec.update(offspring_initial_energy=0.0, reproduction_energy_invested_sum=0.0, reproduction_energy_invested_count=0,
                       parent_energy_after_reproduction_sum=0.0, parent_energy_after_reproduction_count=0)
        if self.coop:
            rec.update(energy_donated=0.0, energy_donated_to_kin=0.0, energy_received=0.0)
        self.stats[agent] = rec

    def _finalize(self, agent, cause, step):
        """the event-log part of `_finalize_agent_record` (MR:1197-1234): a no-op for a record that is already closed"""
        if agent not in self.live_order:
            return
        self.live_order.remove(agent)
        self.completed_order.append(agent)
        rec = self.stats[agent]
        rec["cumulative_reward"] += self.step_reward.get(agent, 0.0)  # MR:1202-1204: the step's reward is added once more
        rec["death_cause"] = cause
        rec["death_step"] = step
        rec["offspring_count"] = self.agent_offspring_counts.get(agent, rec["offspring_count"])
        rec["avg_energy"] = rec["avg_energy_sum"] / max(rec["avg_energy_steps"], 1)
        evt = self.agent_event_log[agent]
        evt["death_step"] = step
        evt["death_cause"] = cause
        self.agent_live_offspring_ids.pop(agent, None)

    def record_order(self):
        return list(self.live_order) + list(self.completed_order)

    def get_all_agent_stats(self):
        """`get_all_agent_stats` (MR:1406-1412): copies of all agent records, live ones first"""
        out = {}
        for a in self.record_order():
            rec = dict(self.stats[a])
            rec["offspring_ids"] = list(rec["offspring_ids"])
            out[a] = rec
        return out

    def get_total_offspring_by_type(self):
        counts = {"predator": 0, "prey": 0}
        for a in self.record_order():
            counts[self.stats[a]["policy_group"]] += self.stats[a]["offspring_count"]
        return counts

    def _donate(self, a, s, gain, order, pos, E, termd):
        """`_apply_cooperative_donation` (COOP:537-589): a cooperation_rate share of a positive gain goes, in equal parts, to the
        live same-species agents within Chebyshev distance cooperation_range; returns what the donor keeps"""
        if not self.coop or self.prev[a][3] is None or gain <= 0.0:
            return gain
        rate = float(self.prev[a][3])
        if rate <= 0.0:
            return gain
        p = pos[a]
        nbrs = [o for o in order[s] if o != a and o not in termd and max(abs(pos[o][0] - p[0]), abs(pos[o][1] - p[1])) <= self.coop_range]
        if not nbrs:
            return gain
        total = rate * gain
        share = total / len(nbrs)
        kin = 0.0
        for o in nbrs:
            E[o] += share
            if o in self.live_order:
                self.stats[o]["energy_received"] += share
            pa, pb = self.agent_parents.get(a), self.agent_parents.get(o)
            if o == pa or a == pb or (pa is not None and pa == pb):  # `_is_kin` (COOP:529-535)
                kin += share
        if a in self.live_order:
            self.stats[a]["energy_donated"] += total
            self.stats[a]["energy_donated_to_kin"] += kin
        self.energy_donated[s] += total
        self.kin_donation[s] += kin
        return gain - total

    def step(self, t, action_dict, rows, state, newborn, grass, time_limit):
        """The reference's own step replayed on the host (positions, energies, who eats whom), checked against what the device
        reports: the survivors' cells and energies, the PPG_ROW_ATE / PPG_ROW_REPRODUCED flags and who is gone."""
        cfg, prev = self.cfg, self.prev
        agents = [a for a in self.prev_agents if a in state] + list(newborn)
        is_pred = lambda a: "predator" in a  # noqa: E731
        order = ([a for a in self.prev_agents if is_pred(a)], [a for a in self.prev_agents if not is_pred(a)])  # *_positions order
        E, deltas, pos = {}, {}, {}
        self.step_reward = {}
        cell = ({}, {})  # grid_world_state layers 0 / 1 (float32): what `_get_move` tests (MR:636)
        # Step 1 (MR:546-574): basal cost, ageing
        for a in self.prev_agents:
            s = 0 if is_pred(a) else 1
            rate = prev[a][3]
            decay = self.loss[s] * (float(rate) if rate is not None else 1.0) if self.scale_loss else self.loss[s]
            if self.cad and rate is not None and self.speed_coeff > 0.0:
                decay *= 1.0 + self.speed_coeff * float(rate)  # CAD:626-633
            E[a] = prev[a][1] - decay
            deltas[a] = {"decay": -decay, "move": 0.0, "eat": 0.0, "repro": 0.0}
            pos[a] = prev[a][0]
            cell[s][pos[a]] = np.float32(E[a])
        termd = set()
        if self.cad:  # CAD:638-650, 972-1000: age caps, after everybody's basal cost
            for a in self.prev_agents:
                limit = self.max_age.get("predator" if is_pred(a) else "prey")
                if isinstance(limit, (int, float)) and limit >= 0 and prev[a][2] + 1 >= limit:
                    termd.add(a)
                    cell[0 if is_pred(a) else 1][pos[a]] = np.float32(0.0)
                    self.agent_event_log[a]["lifecycle_events"].append({"t": int(t), "event": "max_age_reached", "age": int(prev[a][2] + 1)})
                    self.stats[a]["age_expired_step"] = int(t)  # CAD:985-988
                    self.step_reward[a] = 0.0
                    self._finalize(a, "max_age", t)
        # Step 2 (MR:576-584)
        g_now = {c: [name, min(e + self.grass_gain, self.grass_max)] for c, (name, e) in self.prev_grass.items()}
        # Step 3 (MR:586-615): moves in the order of the action dict; a target whose own-species layer reads > 0 blocks
        for a, action in action_dict.items():
            if a not in prev or a in termd:
                continue
            s = 0 if is_pred(a) else 1
            if self.cad:  # CAD:674-681: the move happens once the accumulator reaches 1
                acc = prev[a][4] + self._move_rate(prev[a][3])
                if acc < 1.0:
                    continue
            old = pos[a]
            new = self._target(old, action)
            if cell[s].get(new, 0.0) > 0:
                new = old
            cost = self._cost(s, old, new, prev[a][3])
            E[a] -= cost
            deltas[a]["move"] -= cost
            cell[s][old] = np.float32(0.0)
            cell[s][new] = np.float32(E[a])
            pos[a] = new
            rec = self.stats[a]  # MR:609-614
            rec["avg_energy_sum"] += E[a]
            rec["avg_energy_steps"] += 1
            rec["distance_traveled"] += float(np.linalg.norm(np.array(new) - np.array(old)))
            rec["movement_energy_spent"] += cost
        # Step 4a (MR:274-280): starvation, `agent_energies` order
        for a in self.prev_agents:
            if E[a] <= 0:
                termd.add(a)
                self.step_reward[a] = 0  # MR:716
                self._finalize(a, "starved", t)
        # Step 4b (MR:790-834): prey that are still alive eat the patch they stand on
        for a in order[1]:
            if a in termd:
                continue
            g = g_now.get(pos[a])
            if g is None:
                self._reward(a, _role(cfg.get("reward_prey_step", 0.0), a))
                continue
            if self.cad:  # CAD:909-925: a capped bite, the rest stays on the patch
                gain = min(float(g[1]), self.cap_grass)
                rest = float(g[1]) - gain
            else:
                gain = self._donate(a, 1, float(g[1]), order, pos, E, termd) if self.coop else self._gain(float(g[1]), prev[a][3])
                rest = 0.0
            E[a] += gain
            deltas[a]["eat"] = gain
            g[1] = rest if rest > 0.0 else 0.0
            self._reward(a, _role(cfg.get("reward_prey_eat_grass", 0.0), a))
            self.stats[a]["times_ate"] += 1
            self.stats[a]["energy_gained"] += gain
            if self.cad:
                self.agent_event_log[a]["eating_events"].append({"t": int(t), "id_eaten": g[0], "alive_before_bite": True,
                                                                 "bite_size": float(gain), "energy_after": float(E[a])})
            else:
                self.agent_event_log[a]["eating_events"].append({"t": int(t), "id_eaten": g[0], "energy_after": float(E[a])})
            if not rows.get(a, 0) & ROW_ATE:
                self.inexact_chains += 1
        # Step 4c (MR:720-788): predators, predator_positions order.  The first prey of `agent_positions` on the cell is caught —
        # a prey stays there until Step 5, also one that starved or was caught before in this step
        for a in order[0]:
            if a in termd:
                continue
            if self.cad:  # CAD:837-848: the nearest prey within Chebyshev distance 1, the first of `agent_positions` on ties
                q, best = None, None
                for o in order[1]:
                    d = max(abs(pos[a][0] - pos[o][0]), abs(pos[a][1] - pos[o][1]))
                    if d <= 1 and (best is None or d < best):
                        q, best = o, d
            else:
                q = next((o for o in order[1] if pos[o] == pos[a]), None)
            if q is not None and self.sat_cd >= 0 and t < self.sat_until.get(a, 0):
                q = None  # still digesting (MR:734-740)
            if q is None:
                self._reward(a, _role(cfg.get("reward_predator_step", 0.0), a))
                if rows.get(a, 0) & ROW_ATE:
                    self.inexact_chains += 1
                continue
            pe = float(E[q])
            gain = self._donate(a, 0, pe, order, pos, E, termd) if self.coop else self._gain(min(pe, self.cap_prey), prev[a][3])
            E[a] += gain
            deltas[a]["eat"] = gain
            if self.sat_cd > 0:
                self.sat_until[a] = int(t) + self.sat_cd  # MR:756-757
            self._reward(a, _role(cfg.get("reward_predator_catch_prey", 0.0), a))
            self.stats[a]["times_ate"] += 1
            self.stats[a]["energy_gained"] += gain
            termd.add(q)
            penalty = _role(cfg.get("penalty_prey_caught", 0.0), q)  # MR:765-773
            self.step_reward[q] = penalty
            if q in self.live_order:
                self.stats[q]["cumulative_reward"] += penalty
            self._finalize(q, "eaten", t)
            if self.cad:
                self.agent_event_log[a]["eating_events"].append({"t": int(t), "id_eaten": q, "bite_size": float(pe), "energy_after": float(E[a])})
            else:
                self.agent_event_log[a]["eating_events"].append({"t": int(t), "id_eaten": q, "energy_after": float(E[a])})
            if not rows.get(a, 0) & ROW_ATE:
                self.inexact_chains += 1
        # what the device reports: who is gone, where the survivors stand
        for a in self.prev_agents:
            if (a in termd) != (a not in state) or (a in state and pos[a] != state[a][0]):
                self.inexact_chains += 1
        # Step 6 (MR:836-1010): births, predators first; the k-th newborn row of a species belongs to the k-th parent
        new_agents = list(newborn)
        for s, role in enumerate(("predator", "prey")):
            parents = [a for a in order[s] if a in state and rows.get(a, 0) & ROW_REPRODUCED]
            children = [a for a in new_agents if is_pred(a) == (s == 0)]
            if len(parents) != len(children):
                self.inexact_chains += 1
            for par, child in zip(parents, children):
                self._register(child, par, int(t), state[child][3])
                self.agent_live_offspring_ids[par].append(child)
                self.agent_offspring_counts[par] += 1
                self.agent_event_log[par]["reproduction_events"].append({"t": int(t), "child_id": child})
                ce = self._offspring_energy(par, s, E[par], prev[par][3])
                if ce != state[child][1]:
                    self.inexact_chains += 1
                E[par] -= ce
                deltas[par]["repro"] = -ce
                self.stats[par]["offspring_count"] += 1
                if not self.cad:  # MR:900-908
                    self.stats[par]["reproduction_energy_invested_sum"] += ce
                    self.stats[par]["reproduction_energy_invested_count"] += 1
                    self.stats[par]["parent_energy_after_reproduction_sum"] += E[par]
                    self.stats[par]["parent_energy_after_reproduction_count"] += 1
                    self.stats[child]["offspring_initial_energy"] = ce
                self.step_reward[child] = 0
                r = _role(cfg.get(f"reproduction_reward_{role}", 0.0), par)
                self._reward(par, r)
                self.agent_event_log[par]["reward_events"].append({"t": int(t), "reproduction_reward": float(r),
                                                                   "cumulative_reward": float(self.cumulative_reward[par])})
                deltas[child] = {"decay": 0.0, "move": 0.0, "eat": 0.0, "repro": 0.0}
        for a in self.prev_agents:
            if a in state and E[a] != state[a][1]:
                self.inexact_chains += 1
        step_data = {}
        for a in agents:
            p, e, ag, _, _ = state[a]
            d = deltas[a]
            step_data[a] = {"position": p, "energy": e, "energy_decay": d["decay"], "energy_movement": d["move"],
                            "energy_eating": d["eat"], "energy_reproduction": d["repro"], "age": ag,
                            "offspring_count": self.agent_offspring_counts[a],
                            "offspring_ids": self.agent_live_offspring_ids.get(a, []), "parent": self.agent_parents.get(a)}
        if not self.cad or self.record_step_data:
            self.per_step_agent_data.append(step_data)
        if time_limit:  # MR:453-454: every record still open is closed with the step counter already advanced
            for a in list(self.live_order):
                self._finalize(a, "time_limit", int(t) + 1)
        self.prev, self.prev_agents, self.prev_grass = dict(state), list(agents), dict(grass)

    def _reward(self, agent, r):
        """`self.rewards[agent] = r` plus the running totals of the record and of the reward events"""
        self.step_reward[agent] = r
        self.cumulative_reward[agent] += r
        self.stats[agent]["cumulative_reward"] += r

    def _offspring_energy(self, parent, s, parent_energy, value):
        """MR:897 / 984: the configured initial energy; INV:546-557: the parent's energy x its investment fraction"""
        if self.init_e is not None:
            return self.init_e[s]
        if value is not None:
            fraction = float(value)
        else:
            role = "predator" if s == 0 else "prey"
            fraction = float(self.cfg.get("founder_genome", {}).get(role, {}).get("offspring_investment_fraction_mean", 0.35))
        return float(parent_energy) * fraction

    def export(self, path):
        with open(path, "w", encoding="utf-8") as f:
            json.dump(self.agent_event_log, f, indent=2)
