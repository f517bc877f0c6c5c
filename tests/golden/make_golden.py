#!/usr/bin/env python
"""Generate tests/golden/env_<Name>.npz by running the UNMODIFIED reference
(/root/reference/opfgym) in the build container.

pandapower / simbench / gymnasium are not installable here, so they are stubbed
(see _ref_stubs.py): the stand-in grids replace SimBench, and ``pp.runpp`` is the
CPU oracle (oracle/pf.py).  Every other line that runs -- env construction,
``_sampling`` hooks, ``_apply_actions``, objective, constraints, reward,
``_get_obs`` -- is the reference's own code.  For each benchmark env the script
records, per sample: every per-environment net column after ``reset``, the
action, and the outputs of ``step``.

Run from the repo root:  python tests/golden/make_golden.py [case names]
(needs /root/reference; the committed .npz files are what the tests read).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

N_SAMPLES = 12
PROFILE_STEPS = 672


def cases(ref_envs):
    class LoadSheddingPd3(ref_envs.LoadShedding):
        # pandas >= 3 refuses to write floats into the int column that
        # `net.pwl_cost['cp1_eur_per_mw'] = 0` creates (load_shedding.py:115);
        # cast it once, semantics unchanged.
        def _define_opf(self, *a, **k):
            net, prof = super()._define_opf(*a, **k)
            net.pwl_cost["cp1_eur_per_mw"] = net.pwl_cost["cp1_eur_per_mw"].astype(float)
            return net, prof

    return {
        "VoltageControl": (ref_envs.VoltageControl, {}),
        "QMarket": (ref_envs.QMarket, {}),
        "EcoDispatch": (ref_envs.EcoDispatch, {}),
        "LoadShedding": (LoadSheddingPd3, {}),
        "MaxRenewable": (ref_envs.MaxRenewable, {}),
        # non-default reward / constraint parameters
        "VoltageControl_param": (ref_envs.VoltageControl, dict(
            reward_function="parameterized",
            reward_function_params=dict(valid_reward=0.7, invalid_penalty=0.3,
                                        invalid_objective_share=0.5, penalty_weight=None,
                                        clip_range=(-3.0, 1.0)),
            constraint_params=dict(only_worst_case_violations=True, penalty_power=2.0,
                                   penalty_factor=1.5, violation_count_penalty=0.1),
            voltage_band=0.02, max_loading=25)),
        # incremental actions, objective relative to the reset state, loads observed per bus
        "VoltageControl_diff": (ref_envs.VoltageControl, dict(
            diff_objective=True, diff_action_step_size=0.2, bus_wise_obs=True)),
        "EcoDispatch_replacement": (ref_envs.EcoDispatch, dict(
            reward_function="replacement",
            reward_function_params=dict(valid_reward=0.5, penalty_weight=0.3),
            max_loading=20)),
    }


def main():
    import _ref_stubs
    _ref_stubs.install(profile_steps=PROFILE_STEPS)
    import opfgym.envs as ref_envs

    # pandas >= 3 hands out read-only `.values`; opf_env.py:453-455 scales that array in place
    # (a private copy under the pandas versions the reference targets) -- restore that behaviour
    import pandas as pd
    pd.Series.values = property(lambda self: self.to_numpy().copy())
    only = set(sys.argv[1:])
    for name, (cls, kw) in cases(ref_envs).items():
        if only and name not in only:
            continue
        env = cls(train_data="full_uniform", test_data="full_uniform", seed=7, **kw)
        tables = ("load", "sgen", "storage", "gen", "ext_grid", "poly_cost", "pwl_cost")
        out = {k: [] for k in ("action", "obs", "reward", "valids", "violations", "penalties",
                               "cost", "objective", "vm_pu", "va_degree", "line_loading",
                               "trafo_loading", "ext_p", "ext_q", "reset_obs", "initial_obj")}
        cols = {}
        pre = {}
        original_apply = env._apply_actions

        def recording_apply(action, *a, **kw):
            # snapshot the action columns as the sampler left them (reset applies the
            # centre action right afterwards, opf_env.py:201-207)
            if recording_apply.armed:
                recording_apply.armed = False
                for t, c, _ in env.act_keys:
                    pre.setdefault((t, c), []).append(env.net[t][c].to_numpy(float).copy())
            return original_apply(action, *a, **kw)
        env._apply_actions = recording_apply
        for k in range(N_SAMPLES):
            recording_apply.armed = True
            reset_obs, _ = env.reset(seed=100 + k)
            for t in tables:
                df = env.net[t]
                for c in df.columns:
                    if c == "points":
                        if not len(df):
                            continue
                        pts = np.array([[seg for seg in p] for p in df[c]], dtype=float)
                        cols.setdefault((t, c), []).append(pts.reshape(len(df), -1))
                    elif df[c].dtype.kind in "fiub":
                        cols.setdefault((t, c), []).append(df[c].to_numpy(float).copy())
            out["initial_obj"].append(np.sum(getattr(env, "initial_obj", 0.0)))
            a = env.action_space.sample()
            if k % 4 == 3:
                a = a * 1.3 - 0.15          # exercise the [0,1] clipping
            obs, reward, term, trunc, info = env.step(a)
            assert term and not trunc
            out["reset_obs"].append(reset_obs)
            out["action"].append(a.astype(float))
            out["obs"].append(np.asarray(obs, float))
            out["reward"].append(float(reward))
            out["valids"].append(np.asarray(info["valids"], bool))
            out["violations"].append(np.asarray(info["violations"], float))
            out["penalties"].append(np.asarray(info["unscaled_penalties"], float))
            out["cost"].append(float(info["cost"]))
            out["objective"].append(float(np.sum(env.calculate_objective())))
            out["vm_pu"].append(env.net.res_bus.vm_pu.to_numpy().copy())
            out["va_degree"].append(env.net.res_bus.va_degree.to_numpy())
            out["line_loading"].append(env.net.res_line.loading_percent.to_numpy())
            out["trafo_loading"].append(env.net.res_trafo.loading_percent.to_numpy())
            out["ext_p"].append(env.net.res_ext_grid.p_mw.to_numpy())
            out["ext_q"].append(env.net.res_ext_grid.q_mvar.to_numpy())
        arrays = {f"out/{k}": np.array(v) for k, v in out.items()}
        for (t, c), v in cols.items():
            arrays[f"col/{t}/{c}"] = np.array(v)
        for (t, c), v in pre.items():
            arrays[f"pre/{t}/{c}"] = np.array(v)
        arrays["meta/constraints"] = np.array([type(c).__name__ for c in env.constraints])
        arrays["meta/obs_low"] = env.observation_space.low
        arrays["meta/obs_high"] = env.observation_space.high
        arrays["meta/kwargs"] = np.array(repr(kw))
        path = os.path.join(HERE, f"env_{name}.npz")
        np.savez_compressed(path, **arrays)
        print(name, "obs", env.observation_space.shape[0], "act", env.action_space.shape[0],
              "valid share", np.mean([v.all() for v in out["valids"]]), "->",
              os.path.relpath(path, ROOT), os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
