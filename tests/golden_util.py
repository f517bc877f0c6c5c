"""Replay of the reference-generated fixtures (tests/golden/env_*.npz, produced by
tests/golden/make_golden.py from the UNMODIFIED reference) on a BatchedOpfEnv."""
import ast
import glob
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PROFILE_STEPS = 672


def fixture_names():
    return sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(HERE, "golden", "env_*.npz")))


def load(name):
    z = np.load(os.path.join(HERE, "golden", f"env_{name}.npz"), allow_pickle=False)
    kwargs = ast.literal_eval(str(z["meta/kwargs"]))
    return z, kwargs


def make_env(name, engine_cls=None, obs_dtype="float64", **extra):
    from opfgym_b200 import envs
    z, kwargs = load(name)
    cls = getattr(envs, name.split("_")[0])
    n = int(z["out/action"].shape[0])
    kw = dict(kwargs)
    kw.update(extra)
    if engine_cls is not None:
        kw["engine_cls"] = engine_cls
    env = cls(num_envs=n, train_data="full_uniform", test_data="full_uniform", seed=7,
              n_profile_steps=PROFILE_STEPS, obs_dtype=obs_dtype, **kw)
    return env, z


def fixture_column(z, table, column):
    """[n_samples, n_rows] values of a per-environment column as the reference had them
    after reset; the LoadShedding pwl segment prices live in ``points``."""
    if table == "pwl_cost" and column in ("price_charge", "price_discharge"):
        pts = z["col/pwl_cost/points"]
        n_rows = pts.shape[1]
        pts = pts.reshape(pts.shape[0], n_rows, -1, 3)
        return pts[:, :, 0 if column == "price_charge" else 1, 2]
    return z[f"col/{table}/{column}"]


def inject(env, z, only=None, pre_action=False):
    """``pre_action``: action columns take the values the reference's sampler left
    (before reset applied the centre action)."""
    t = env.xp
    for (table, column), (start, n_rows) in env.program.layout.columns.items():
        if table.startswith("res_") or n_rows == 0:
            continue
        if only is not None and (table, column) not in only:
            continue
        v = fixture_column(z, table, column)
        if pre_action and f"pre/{table}/{column}" in z.files:
            v = z[f"pre/{table}/{column}"]
        env.col(table, column).copy_(t.as_tensor(np.ascontiguousarray(v), device=env.device))


def to_np(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


def replay_step(env, z):
    """Inject the reference's post-reset state and its action, run ONE engine step
    (no sampling), return the engine outputs as numpy."""
    inject(env, z)
    e = env.engine
    if getattr(env, "diff_objective", False):
        e.enable_objective_offset().copy_(env.xp.as_tensor(z["out/initial_obj"], device=env.device))
    e.actions.copy_(env.xp.as_tensor(z["out/action"], device=env.device))
    e.step()
    if env.device.type == "cuda":
        env.xp.cuda.synchronize()
    nc = len(env.constraints)
    return dict(obs=to_np(e.obs).astype(float), reward=to_np(e.reward), cost=to_np(e.cost),
                objective=to_np(e.objective), valids=to_np(e.valids)[:, :nc].astype(bool),
                violations=to_np(e.violations)[:, :nc], penalties=to_np(e.penalties)[:, :nc],
                vm=to_np(e.column("res_bus", "vm_pu")),
                line_loading=to_np(e.column("res_line", "loading_percent")),
                trafo_loading=to_np(e.column("res_trafo", "loading_percent")),
                converged=to_np(e.converged))


def assert_matches(got, z, reward_rtol=1e-6):
    assert got["converged"].all()
    np.testing.assert_array_equal(got["valids"], z["out/valids"])
    np.testing.assert_allclose(got["vm"], z["out/vm_pu"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(got["line_loading"], z["out/line_loading"], rtol=0, atol=1e-4)
    np.testing.assert_allclose(got["trafo_loading"], z["out/trafo_loading"], rtol=0, atol=1e-4)
    np.testing.assert_allclose(got["reward"], z["out/reward"], rtol=reward_rtol, atol=1e-9)
    initial = z["out/initial_obj"] if "out/initial_obj" in z.files else 0.0   # diff_objective cases
    np.testing.assert_allclose(got["objective"], z["out/objective"] - initial, rtol=reward_rtol, atol=1e-9)
    np.testing.assert_allclose(got["cost"], z["out/cost"], rtol=reward_rtol, atol=1e-9)
    np.testing.assert_allclose(got["violations"], z["out/violations"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(got["penalties"], z["out/penalties"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(got["obs"], z["out/obs"], rtol=1e-6, atol=1e-9)
