"""Branch coverage of the oracle for the paths no reference fixture reaches (SURVEY.md section 4 gaps):
each test states the reference lines the expected value is read from."""
import numpy as np
import pytest


def f32(x):
    return np.float32(x)


def test_cartpole_constants(oracle):
    """cartpole.rs:45-56: f32 constant arithmetic, checked through its observable effects."""
    # theta_threshold = 12*2*PI/360 in f32 = 0x3e567750; strict '>' (cartpole.rs:291-294)
    thr = np.array([0x3e567750], dtype=np.uint32).view(np.float32)[0]
    assert thr == f32(f32(f32(12.0) * f32(2.0)) * f32(np.pi)) / f32(360.0)
    for theta, expect_done in [(thr, False), (np.nextafter(thr, f32(1)), True), (-thr, False),
                               (np.nextafter(-thr, f32(-1)), True)]:
        e = oracle.ScalarEnv(oracle.CARTPOLE)
        e.reset()
        # theta_dot = 0: the Euler step leaves theta unchanged (theta + tau*0)
        e.set_state([0.0, 0.0, theta, 0.0])
        _, _, done, _ = e.step(1)
        assert done == expect_done, (theta, done)
    # x_threshold 2.4, strict
    for x, expect_done in [(f32(2.4), False), (np.nextafter(f32(2.4), f32(3)), True), (f32(-2.4), False)]:
        e = oracle.ScalarEnv(oracle.CARTPOLE)
        e.reset()
        e.set_state([x, 0.0, 0.0, 0.0])
        _, _, done, _ = e.step(1 if x < 0 else 0)
        assert done == expect_done


def test_cartpole_one_step_by_hand(oracle):
    """cartpole.rs:258-277 evaluated with numpy float32 in the same operator order."""
    rng = np.random.default_rng(0)
    for _ in range(200):
        s = rng.uniform(-0.2, 0.2, 4).astype(np.float32)
        a = int(rng.integers(0, 2))
        e = oracle.ScalarEnv(oracle.CARTPOLE)
        e.reset()
        e.set_state(s)
        obs, _, _, _ = e.step(a)
        x, xd, th, thd = s
        g, mp, tm, ln, pml, tau = f32(9.8), f32(0.1), f32(0.1) + f32(1.0), f32(0.5), f32(0.1) * f32(0.5), f32(0.02)
        force = f32(10.0) if a else f32(-10.0)
        c, sn = f32(oracle.cosf(th)), f32(oracle.sinf(th))
        temp = (force + pml * thd * thd * sn) / tm
        thacc = (g * sn - c * temp) / (ln * (f32(4.0) / f32(3.0) - mp * c * c / tm))
        xacc = temp - pml * thacc * c / tm
        want = np.array([x + tau * xd, xd + tau * xacc, th + tau * thd, thd + tau * thacc], dtype=np.float32)
        assert obs.view(np.uint32).tolist() == want.view(np.uint32).tolist()


def test_cartpole_truncation_quirk(oracle):
    """cartpole.rs:296-306: at steps_since_reset >= 500 the step returns reward 1.0, done=false,
    truncated=true, even if the pole has fallen and even with sutton_barto_reward; and it keeps doing so
    until reset() (the counter is only cleared there, :243)."""
    for sb in (0, 1):
        env = oracle.ScalarEnv(oracle.CARTPOLE, sutton_barto_reward=sb)
        env.reset()
        env.steps.value = 498
        env.set_state([0, 0, 0, 0])
        _, r, done, trunc = env.step(0)
        assert (r, done, trunc) == (0.0 if sb else 1.0, False, False)
        env.set_state([3.0, 0, 0.5, 0])            # far outside both thresholds
        _, r, done, trunc = env.step(0)
        assert (r, done, trunc) == (1.0, False, True)
        assert env.sbt.value == 1                   # Some(0), :299
        _, r, done, trunc = env.step(0)
        assert (r, done, trunc) == (1.0, False, True)   # still truncating
        env.reset()
        assert env.steps.value == 0 and env.sbt.value == 0


def test_cartpole_post_termination_rewards(oracle):
    """cartpole.rs:310-347: reward 1 on the terminating step, then 0 (and -1 / -1 with sutton_barto)."""
    for sb, first, later, alive in [(0, 1.0, 0.0, 1.0), (1, -1.0, -1.0, 0.0)]:
        env = oracle.ScalarEnv(oracle.CARTPOLE, sutton_barto_reward=sb)
        env.reset()
        env.set_state([0, 0, 0, 0])
        assert env.step(1)[1] == alive
        env.set_state([2.5, 0, 0, 0])
        _, r, done, trunc = env.step(1)
        assert (r, done, trunc) == (first, True, False) and env.sbt.value == 1
        for k in range(3):
            _, r, done, trunc = env.step(1)
            assert (r, done, trunc) == (later, True, False) and env.sbt.value == 2 + k


def test_cartpole_fresh_env_is_already_terminated(oracle):
    """cartpole.rs:81: a never-reset env holds steps_beyond_terminated = Some(0), so a terminating first step
    takes the :330-347 branch (reward 0.0)."""
    env = oracle.ScalarEnv(oracle.CARTPOLE)
    env.set_state([2.5, 0, 0, 0])
    _, r, done, _ = env.step(1)
    assert (r, done) == (0.0, True)


def test_cartpole_non_euler_branch(oracle):
    """cartpole.rs:278-283 verbatim: x is not advanced, theta_dot is advanced twice."""
    s = np.array([0.1, 0.2, 0.05, -0.3], dtype=np.float32)
    env = oracle.ScalarEnv(oracle.CARTPOLE, is_euler=0)
    env.reset()
    env.set_state(s)
    obs, _, _, _ = env.step(1)
    x, xd, th, thd = s
    g, mp, tm, ln, pml, tau = f32(9.8), f32(0.1), f32(0.1) + f32(1.0), f32(0.5), f32(0.1) * f32(0.5), f32(0.02)
    c, sn = f32(oracle.cosf(th)), f32(oracle.sinf(th))
    temp = (f32(10.0) + pml * thd * thd * sn) / tm
    thacc = (g * sn - c * temp) / (ln * (f32(4.0) / f32(3.0) - mp * c * c / tm))
    xacc = temp - pml * thacc * c / tm
    half = f32(0.5) * tau
    xd2 = xd + half * (xacc + temp)
    thd1 = thd + half * (thacc + temp)
    th2 = th + (tau * thd1 + half * tau * thacc)
    thd2 = thd1 + half * (thacc + temp)
    want = np.array([x, xd2, th2, thd2], dtype=np.float32)
    assert obs.view(np.uint32).tolist() == want.view(np.uint32).tolist()
    assert obs[0] == s[0]


def test_mountain_car_rules(oracle):
    """mountain_car.rs:301-318 by hand, the left-wall rule (:311-313) and goal termination (:318)."""
    rng = np.random.default_rng(1)
    for _ in range(300):
        p, v = f32(rng.uniform(-1.2, 0.6)), f32(rng.uniform(-0.07, 0.07))
        a = int(rng.integers(0, 3))
        env = oracle.ScalarEnv(oracle.MOUNTAIN_CAR)
        env.set_state([p, v])
        obs, r, done, trunc = env.step(a)
        nv = v + ((f32(a) - f32(1.0)) * f32(0.001) + f32(oracle.cosf(f32(3.0) * p)) * f32(-0.0025))
        nv = min(max(nv, f32(-0.07)), f32(0.07))
        np_ = min(max(p + nv, f32(-1.2)), f32(0.6))
        if np_ == f32(-1.2) and nv < 0:
            nv = f32(0)
        assert obs.view(np.uint32).tolist() == np.array([np_, nv], np.float32).view(np.uint32).tolist()
        assert r == -1.0 and not trunc and done == bool(np_ >= f32(0.5) and nv >= 0)
    env = oracle.ScalarEnv(oracle.MOUNTAIN_CAR)
    env.set_state([-1.2, -0.05])
    obs, _, _, _ = env.step(0)
    assert obs[0] == f32(-1.2) and obs[1] == 0.0
    env.set_state([0.499, 0.07])
    assert env.step(2)[2] is True
    hi = oracle.ScalarEnv(oracle.MOUNTAIN_CAR, goal_velocity=0.08)   # unreachable goal velocity
    hi.set_state([0.499, 0.07])
    assert hi.step(2)[2] is False
    # the reference never truncates (:328); an explicit limit is this repo's option
    env = oracle.ScalarEnv(oracle.MOUNTAIN_CAR)
    env.reset()
    assert not any(env.step(1)[3] for _ in range(1000))
    lim = oracle.ScalarEnv(oracle.MOUNTAIN_CAR, max_episode_steps=200)
    lim.reset()
    flags = [lim.step(1)[3] for _ in range(200)]
    assert flags[:199] == [False] * 199 and flags[199] is True


# --- envs the reference lacks: cross-check the C restatement against a float64 transcription of the
# --- Gymnasium equations written independently here (parity unpinned: no reference code exists) ------
def gym_pendulum_f64(th, thdot, u):
    u = float(np.clip(u, -2.0, 2.0))
    an = ((th + np.pi) % (2 * np.pi)) - np.pi
    cost = an ** 2 + 0.1 * thdot ** 2 + 0.001 * u ** 2
    nthdot = float(np.clip(thdot + (3 * 10.0 / 2 * np.sin(th) + 3.0 * u) * 0.05, -8, 8))
    return th + nthdot * 0.05, nthdot, -cost


def gym_acrobot_f64(s, a):
    def dsdt(y):
        t1, t2, d1_, d2_ = y
        m1 = m2 = l1 = 1.0
        lc1 = lc2 = 0.5
        I1 = I2 = 1.0
        g = 9.8
        d1 = m1 * lc1 ** 2 + m2 * (l1 ** 2 + lc2 ** 2 + 2 * l1 * lc2 * np.cos(t2)) + I1 + I2
        d2 = m2 * (lc2 ** 2 + l1 * lc2 * np.cos(t2)) + I2
        phi2 = m2 * lc2 * g * np.cos(t1 + t2 - np.pi / 2.0)
        phi1 = (-m2 * l1 * lc2 * d2_ ** 2 * np.sin(t2) - 2 * m2 * l1 * lc2 * d2_ * d1_ * np.sin(t2)
                + (m1 * lc1 + m2 * l1) * g * np.cos(t1 - np.pi / 2) + phi2)
        dd2 = (a + d2 / d1 * phi1 - m2 * l1 * lc2 * d1_ ** 2 * np.sin(t2) - phi2) / (m2 * lc2 ** 2 + I2 - d2 ** 2 / d1)
        dd1 = -(d2 * dd2 + phi1) / d1
        return np.array([d1_, d2_, dd1, dd2])

    y0 = np.asarray(s, dtype=np.float64)
    dt = 0.2
    k1 = dsdt(y0)
    k2 = dsdt(y0 + dt / 2 * k1)
    k3 = dsdt(y0 + dt / 2 * k2)
    k4 = dsdt(y0 + dt * k3)
    y = y0 + dt / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)

    def wrap(x, m, M):
        while x > M:
            x -= M - m
        while x < m:
            x += M - m
        return x

    y[0], y[1] = wrap(y[0], -np.pi, np.pi), wrap(y[1], -np.pi, np.pi)
    y[2], y[3] = np.clip(y[2], -4 * np.pi, 4 * np.pi), np.clip(y[3], -9 * np.pi, 9 * np.pi)
    term = bool(-np.cos(y[0]) - np.cos(y[1] + y[0]) > 1.0)
    return y, (0.0 if term else -1.0), term


def test_pendulum_against_float64_gymnasium_equations(oracle):
    rng = np.random.default_rng(2)
    for _ in range(500):
        th, thd, u = rng.uniform(-8, 8), rng.uniform(-8, 8), rng.uniform(-2.5, 2.5)
        env = oracle.ScalarEnv(oracle.PENDULUM)
        env.set_state([th, thd])
        obs, r, done, trunc = env.step(u)
        nth, nthd, wr = gym_pendulum_f64(float(f32(th)), float(f32(thd)), float(f32(u)))
        assert np.allclose(obs, [np.cos(nth), np.sin(nth), nthd], atol=2e-5)
        assert r == pytest.approx(wr, rel=1e-5, abs=1e-5) and not done and not trunc
    env = oracle.ScalarEnv(oracle.PENDULUM)
    env.reset()
    flags = [env.step(0.0)[3] for _ in range(200)]
    assert flags[:199] == [False] * 199 and flags[199] is True   # TimeLimit 200


def test_acrobot_against_float64_gymnasium_equations(oracle):
    rng = np.random.default_rng(3)
    n_term = 0
    for _ in range(500):
        s = np.array([rng.uniform(-np.pi, np.pi), rng.uniform(-np.pi, np.pi), rng.uniform(-6, 6), rng.uniform(-10, 10)])
        a = int(rng.integers(0, 3))
        env = oracle.ScalarEnv(oracle.ACROBOT)
        env.set_state(s)
        obs, r, done, trunc = env.step(a)
        y, wr, wterm = gym_acrobot_f64(s.astype(np.float32).astype(np.float64), a - 1.0)
        margin = abs(-np.cos(y[0]) - np.cos(y[1] + y[0]) - 1.0)
        if margin > 1e-3:
            assert done == wterm and r == wr
        if min(np.pi - abs(y[0]), np.pi - abs(y[1])) > 1e-3:   # away from the wrap seam
            want = [np.cos(y[0]), np.sin(y[0]), np.cos(y[1]), np.sin(y[1]), y[2], y[3]]
            assert np.allclose(obs, want, atol=2e-4, rtol=2e-5), (obs, want)
        n_term += int(done)
    assert n_term > 10


def test_mountain_car_continuous_rules(oracle):
    rng = np.random.default_rng(4)
    for _ in range(300):
        p, v, a = f32(rng.uniform(-1.2, 0.6)), f32(rng.uniform(-0.07, 0.07)), f32(rng.uniform(-1.5, 1.5))
        env = oracle.ScalarEnv(oracle.MOUNTAIN_CAR_CONTINUOUS)
        env.set_state([p, v])
        obs, r, done, trunc = env.step(a)
        force = min(max(float(a), -1.0), 1.0)
        nv = float(v) + force * 0.0015 - 0.0025 * np.cos(3 * float(p))
        nv = min(max(nv, -0.07), 0.07)
        np_ = min(max(float(p) + nv, -1.2), 0.6)
        if np_ == -1.2 and nv < 0:
            nv = 0.0
        assert np.allclose(obs, [np_, nv], atol=1e-6)
        if abs(np_ - 0.45) > 1e-5 and abs(nv) > 1e-6:
            term = np_ >= 0.45 and nv >= 0
            assert done == term
            assert r == pytest.approx((100.0 if term else 0.0) - 0.1 * float(a) ** 2, abs=1e-5)


def test_philox_known_answers(oracle):
    """Random123 kat_vectors for philox4x32-10."""
    assert oracle.philox([0, 0, 0, 0], [0, 0]).tolist() == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox([0xffffffff] * 4, [0xffffffff] * 2).tolist() == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]).tolist() == [
        0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_reset_distributions(oracle):
    """cartpole.rs:240 U[-0.05,0.05); mountain_car.rs:281-283 U[-0.6,-0.4), v=0."""
    cp = np.stack([oracle.reset_state(oracle.CARTPOLE, 7, g, 0, 1) for g in range(4000)])
    assert cp.min() >= -0.05 and cp.max() <= 0.05 and abs(cp.mean()) < 2e-3
    mc = np.stack([oracle.reset_state(oracle.MOUNTAIN_CAR, 7, g, 0, 1) for g in range(2000)])
    assert mc[:, 0].min() >= -0.6 and mc[:, 0].max() <= -0.4 and (mc[:, 1] == 0).all()
    # counter-based: independent of call order, distinct per (env, step, tag)
    a = oracle.reset_state(0, 7, 5, 3, 0)
    assert (a == oracle.reset_state(0, 7, 5, 3, 0)).all()
    assert not (a == oracle.reset_state(0, 7, 5, 3, 1)).all() and not (a == oracle.reset_state(0, 7, 5, 4, 0)).all()
