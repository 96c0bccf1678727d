"""GPU parity of the domain-randomisation noise lambdas (reference tasks/base/vec_task.py:562-618, applied at
:314-315 / :338-339) -- the kernel with the draws supplied against the restated lambda, the Philox path against its own dump,
and the ``randomize: True`` KickEnv path (frequency gate, schedule scaling, action + observation noise)."""
import pytest
import torch

from bez_isaacgym_b200 import bez_model as bm
from tests import _util as U

pytestmark = pytest.mark.gpu


def _ops():
    from bez_isaacgym_b200 import ops
    return ops


@pytest.mark.parametrize("dist,op,p0,p1,c0,c1", [("gaussian", "additive", 0.0, 0.002, 0.0, 0.0), ("gaussian", "additive", 0.1, 0.02, -0.05, 0.3),
                                                 ("gaussian", "scaling", 1.0, 0.05, 1.0, 0.01), ("uniform", "additive", -0.01, 0.02, 0.0, 0.0),
                                                 ("uniform", "scaling", 0.9, 1.1, 0.95, 1.05)])
@pytest.mark.parametrize("shape", [(4096, 54), (4099, 18), (1, 3), (7,)])
def test_dr_noise_matches_reference_lambda(dist, op, p0, p1, c0, c1, shape):
    from oracle import rl_games_oracle as rg
    ops = _ops()
    g = torch.Generator().manual_seed(sum(shape))
    x, corr = torch.randn(shape, generator=g), torch.randn(shape, generator=g)
    white = torch.randn(shape, generator=g) if dist == "gaussian" else torch.rand(shape, generator=g)
    want = rg.dr_noise_lambda(x, corr, white, dist, op, p0, p1, c0, c1)
    if dist == "gaussian":
        cfg = ops.make_noise_cfg(dist, op, a=p1, b=p0, a_corr=c1, b_corr=c0)
    else:
        cfg = ops.make_noise_cfg(dist, op, a=p1 - p0, b=p0, a_corr=c1 - c0, b_corr=c0)
    got = ops.dr_noise(x.cuda(), cfg, corr=corr.cuda(), white=white.cuda(), out=torch.empty(shape, device="cuda"))
    U.assert_close(got, want, what="noise lambda")
    # in place, and without a correlated term
    xin = x.cuda().clone()
    ops.dr_noise(xin, cfg, corr=corr.cuda(), white=white.cuda())
    assert torch.equal(xin, got)
    got0 = ops.dr_noise(x.cuda(), cfg, corr=None, white=white.cuda(), out=torch.empty(shape, device="cuda"))
    U.assert_close(got0, rg.dr_noise_lambda(x, torch.zeros(shape), white, dist, op, p0, p1, c0, c1), what="no corr")


@pytest.mark.parametrize("dist", ["gaussian", "uniform"])
def test_dr_noise_philox_path_uses_the_dumped_draws(dist):
    ops = _ops()
    n = 262144 * 18 + 3
    x = torch.randn(n, device="cuda")
    cfg = ops.make_noise_cfg(dist, "additive", a=0.5, b=0.25)
    w = ops.dr_fill(9, 77, torch.empty(n, device="cuda"), dist)
    got = ops.dr_noise(x, cfg, seed=9, step=77, out=torch.empty_like(x))
    want = ops.dr_noise(x, cfg, white=w, out=torch.empty_like(x))
    assert torch.equal(got, want)
    wd = w.double()
    if dist == "gaussian":
        assert abs(float(wd.mean())) < 2e-3 and abs(float(wd.var()) - 1.0) < 5e-3
    else:
        assert float(w.min()) >= 0.0 and float(w.max()) < 1.0 and abs(float(wd.mean()) - 0.5) < 1e-3
    assert not torch.equal(w, ops.dr_fill(9, 78, torch.empty(n, device="cuda"), dist))


def test_kickenv_with_randomize_applies_action_and_observation_noise():
    """``task.randomize: True`` with the yaml's observation / action noise (cfg/task/bez_kick.yaml:153-162): the step equals
    the noise-free step on the noised actions, plus the observation noise -- with the draws read back through bezk_dr_fill."""
    from bez_isaacgym_b200.synthetic_sim import SyntheticGym
    from bez_isaacgym_b200.tasks import KickEnv
    ops = _ops()
    n = 2048
    params = {"frequency": 3,
              "observations": {"range": [0, .002], "operation": "additive", "distribution": "gaussian"},
              "actions": {"range": [0., .02], "range_correlated": [0.0, 0.01], "operation": "additive", "distribution": "gaussian",
                          "schedule": "linear", "schedule_steps": 4}}

    def make(randomize):
        cfg = bm.default_task_cfg(n)
        cfg["task"] = {"randomize": randomize, "randomization_params": params}
        return KickEnv(cfg, "cuda:0", 0, True, sim=SyntheticGym(n, device="cuda:0", seed=5))

    env, plain = make(True), make(False)
    assert set(env.dr_randomizations) == {"observations", "actions"} and plain.dr_randomizations == {}
    a = env.dr_randomizations["actions"]
    assert a["var"] == 0.0 and a["var_corr"] == 0.0                     # linear schedule at frame 0: scaled to nothing
    acts = torch.randn(n, 18, device="cuda").clamp_(-1, 1)
    env.sim.frame = plain.sim.frame = 1                                  # simulate() makes it 2: 2 - 0 < frequency 3 -> gate closed
    env.reset_buf.fill_(1); plain.reset_buf.fill_(1)                     # resets pending -> apply_randomizations runs
    o1, *_ = env.step(acts)
    plain.step(acts)
    assert env.dr_randomizations["actions"]["var"] == 0.0                # gate closed: parameters unchanged
    env.sim.frame = plain.sim.frame = 5
    env.reset_buf.fill_(1); plain.reset_buf.fill_(1)
    period_before = env._dr_period
    o2, *_ = env.step(acts)
    plain.step(acts)
    assert env._dr_period == period_before + 2                           # both lambdas rebuilt (5 - 0 >= 3)
    a = env.dr_randomizations["actions"]
    assert a["var"] == pytest.approx(0.02) and a["var_corr"] == pytest.approx(0.01)      # schedule saturated: min(6, 4) / 4 = 1
    # third step: reproduce it by hand from the dumped draws
    prm_o = env.dr_randomizations["observations"]
    env.reset_buf.zero_(); plain.reset_buf.zero_()
    plain.progress_buf.copy_(env.progress_buf); plain.dof_state.copy_(env.dof_state); plain.root_states.copy_(env.root_states)
    plain.net_contact.copy_(env.net_contact)
    plain._rng_step = env._rng_step
    calls_a, calls_o = a.get("calls", 0), prm_o.get("calls", 0)
    sid_a, sid_o = (env._dr_period << 1) | 1, ((env._dr_period - 1) << 1) | 0
    o3, r3, d3, _ = env.step(acts)
    corr_a = a["corr"]
    w_a = ops.dr_fill(env._dr_seed, (sid_a << 32) + calls_a + 1, torch.empty(n, 18, device="cuda"))
    noised = acts + ((corr_a * 0.01 + 0.0) + w_a * 0.02 + 0.0)
    p3, pr3, pd3, _ = plain.step(noised)
    assert torch.equal(d3, pd3)
    torch.testing.assert_close(r3, pr3, rtol=1e-5, atol=1e-6)
    w_o = ops.dr_fill(env._dr_seed, (sid_o << 32) + calls_o + 1, torch.empty(n, 54, device="cuda"))
    corr_o = prm_o["corr"]
    want = p3["obs"] + ((corr_o * 0.0 + 0.0) + w_o * 0.002 + 0.0)
    torch.testing.assert_close(o3["obs"], want, rtol=1e-5, atol=1e-6)
    assert float((o3["obs"] - p3["obs"]).abs().max()) > 1e-4            # the noise is really there


def test_physx_randomisation_keys_are_refused():
    from bez_isaacgym_b200.tasks import KickEnv
    cfg = bm.default_task_cfg(64)
    cfg["task"] = {"randomize": True, "randomization_params": {"sim_params": {"gravity": {"range": [0, 0.4]}}}}
    with pytest.raises(NotImplementedError):
        KickEnv(cfg, "cuda:0", 0, True)


def test_observation_noise_with_finite_clip_obs_refreshes_the_clipped_copy_in_the_same_pass():
    """vec_task.py:338-343: the clamp follows the noise.  With ``clipObservations`` finite the noise kernel also writes
    ``clamp(noised, -clip, clip)`` (``bezk_dr_noise_clip``): what ``step`` returns equals clamp(obs_buf) bit for bit."""
    from bez_isaacgym_b200.synthetic_sim import SyntheticGym
    from bez_isaacgym_b200.tasks import KickEnv
    ops = _ops()
    n = 1023                                                            # odd size: the ragged-tail path of the kernel
    cfg = bm.default_task_cfg(n)
    cfg["env"]["clipObservations"] = 1.5
    cfg["task"] = {"randomize": True, "randomization_params": {
        "frequency": 1, "observations": {"range": [0, .5], "operation": "additive", "distribution": "gaussian"}}}
    env = KickEnv(cfg, "cuda:0", 0, True, sim=SyntheticGym(n, device="cuda:0", seed=9))
    out, *_ = env.step(torch.zeros(n, 18, device="cuda"))
    got = out["obs"]
    assert got.data_ptr() == env.obs_clipped_buf.data_ptr()
    assert torch.equal(torch.nan_to_num(got, nan=7.0), torch.nan_to_num(torch.clamp(env.obs_buf, -1.5, 1.5), nan=7.0))
    assert float(got.abs().max()) <= 1.5 and float(env.obs_buf.abs().max()) > 1.5
    # the stand-alone entry: y and its clipped copy from one launch, ragged total
    x = torch.randn(4099, device="cuda") * 3
    y, yc = torch.empty_like(x), torch.empty_like(x)
    kc = ops.make_noise_cfg("uniform", "scaling", a=0.5, b=0.75)
    ops.dr_noise(x, kc, seed=3, step=4, out=y, out_clipped=yc, clip=2.0)
    y2 = ops.dr_noise(x, kc, seed=3, step=4, out=torch.empty_like(x))
    assert torch.equal(y, y2) and torch.equal(yc, torch.clamp(y, -2.0, 2.0))
