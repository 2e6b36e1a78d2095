"""Float64 transcriptions of the five classic-control envs, independent of oracle/mgym_oracle.c and of the CUDA code:
MountainCarContinuous-v0, Pendulum-v1 and Acrobot-v1 -- which the reference does NOT implement
(`/root/reference/src/classic_control.rs:1-2` declares only cartpole and mountain_car) -- from Gymnasium's published
equations, and CartPole-v1 / MountainCar-v0 from the reference's own (`cartpole.rs:251-348`,
`mountain_car.rs:293-330`).  Plain Python/numpy float64, one env, no batching.

TEST INFRASTRUCTURE: this is the yardstick the f32 oracle and the device path are measured against for the kinds
whose parity is otherwise unpinned.  It is not an implementation of the product path.

The teacher-forced trace protocol follows the reference's own known-answer harness (`src/testing.rs:65-134`):
every step starts from the fixture's recorded state, so the comparison never drifts however chaotic the system is.
"""
import math

import numpy as np

CARTPOLE, MOUNTAIN_CAR, MOUNTAIN_CAR_CONTINUOUS, PENDULUM, ACROBOT = 0, 1, 2, 3, 4
# CartPole's 500 is the reference's own early return (cartpole.rs:296-306); MountainCar-v0 never truncates in the
# reference (mountain_car.rs:328); the other three are Gymnasium's TimeLimit wrappers
TIME_LIMIT = {CARTPOLE: 500, MOUNTAIN_CAR: 0, MOUNTAIN_CAR_CONTINUOUS: 999, PENDULUM: 200, ACROBOT: 500}
STATE_DIM = {CARTPOLE: 4, MOUNTAIN_CAR: 2, MOUNTAIN_CAR_CONTINUOUS: 2, PENDULUM: 2, ACROBOT: 4}
OBS_DIM = {CARTPOLE: 4, MOUNTAIN_CAR: 2, MOUNTAIN_CAR_CONTINUOUS: 2, PENDULUM: 3, ACROBOT: 6}


# ---- CartPole-v1 and MountainCar-v0: the two kinds the reference DOES implement, transcribed a third time (after
# ---- oracle/mgym_oracle.c and the CUDA code) in float64, from the reference's equations, so that long traces reach
# ---- what its 100-step fixtures do not: the 500-step truncation, the wall and the goal ------------------------------
def cartpole_step(state, action, count):
    """/root/reference/src/classic_control/cartpole.rs:251-348, Euler integrator, default rewards, first step after
    reset() (steps_beyond_terminated = None).  -> (next_state, obs, reward, terminated, truncated, seam)"""
    gravity, masscart, masspole, length, force_mag, tau = 9.8, 1.0, 0.1, 0.5, 10.0, 0.02
    total_mass, polemass_length = masspole + masscart, masspole * length
    theta_threshold, x_threshold = 12 * 2 * math.pi / 360, 2.4
    x, x_dot, theta, theta_dot = (float(v) for v in state)
    force = force_mag if int(action) == 1 else -force_mag                                  # :258-262
    costheta, sintheta = math.cos(theta), math.sin(theta)                                  # :264-265
    temp = (force + polemass_length * theta_dot * theta_dot * sintheta) / total_mass       # :267-268
    thetaacc = (gravity * sintheta - costheta * temp) / (
        length * (4.0 / 3.0 - masspole * costheta * costheta / total_mass))                # :269-270
    xacc = temp - polemass_length * thetaacc * costheta / total_mass                       # :271
    x, x_dot = x + tau * x_dot, x_dot + tau * xacc                                         # :274-275
    theta, theta_dot = theta + tau * theta_dot, theta_dot + tau * thetaacc                 # :276-277
    terminated = bool(x < -x_threshold or x > x_threshold or theta < -theta_threshold or theta > theta_threshold)
    seam = min(abs(abs(x) - x_threshold), abs(abs(theta) - theta_threshold)) < 1e-6
    nxt = np.array([x, x_dot, theta, theta_dot])
    if count + 1 >= 500:  # :296-306: the early return -- truncated, NOT done, reward 1, whatever the pole did
        return nxt, nxt.copy(), 1.0, False, True, False
    return nxt, nxt.copy(), 1.0, terminated, False, seam                                    # :310-329


def cartpole_reset(rng):
    return rng.uniform(-0.05, 0.05, size=4)                                                # :240


def mountain_car_step(state, action, goal_velocity=0.0):
    """/root/reference/src/classic_control/mountain_car.rs:293-330"""
    min_position, max_position, max_speed, goal_position, force, gravity = -1.2, 0.6, 0.07, 0.5, 0.001, 0.0025
    position, velocity = float(state[0]), float(state[1])
    velocity += (int(action) - 1) * force + math.cos(3 * position) * (-gravity)            # :301-302
    seam = abs(abs(velocity) - max_speed) < 1e-8
    velocity = min(max(velocity, -max_speed), max_speed)                                   # :304
    position += velocity                                                                   # :306
    seam = seam or abs(position - min_position) < 1e-6 or abs(position - max_position) < 1e-6
    position = min(max(position, min_position), max_position)                              # :308
    if position == min_position and velocity < 0:                                          # :311-313
        velocity = 0.0
    terminated = bool(position >= goal_position and velocity >= goal_velocity)             # :318
    seam = seam or abs(position - goal_position) < 1e-6
    nxt = np.array([position, velocity])
    return nxt, nxt.copy(), -1.0, terminated, seam                                          # :319, :328


def mountain_car_reset(rng):
    return np.array([rng.uniform(-0.6, -0.4), 0.0])                                        # :281-285


# ---- MountainCarContinuous-v0 (gymnasium/envs/classic_control/continuous_mountain_car.py) -------------------------
def mountain_car_continuous_step(state, action, goal_velocity=0.0):
    """-> (next_state, obs, reward, terminated, seam).  `seam`: the outcome depends on a comparison the f64 and f32
    evaluations may legitimately decide differently (a value within 1e-6 of a clamp bound or the goal)."""
    min_action, max_action = -1.0, 1.0
    min_position, max_position, max_speed = -1.2, 0.6, 0.07
    goal_position, power = 0.45, 0.0015
    position, velocity = float(state[0]), float(state[1])
    force = min(max(float(action), min_action), max_action)
    velocity += force * power - 0.0025 * math.cos(3 * position)
    seam = abs(abs(velocity) - max_speed) < 1e-7
    if velocity > max_speed:
        velocity = max_speed
    if velocity < -max_speed:
        velocity = -max_speed
    position += velocity
    seam = seam or abs(position - min_position) < 1e-6 or abs(position - max_position) < 1e-6
    if position > max_position:
        position = max_position
    if position < min_position:
        position = min_position
    if position == min_position and velocity < 0:
        velocity = 0.0
    terminated = bool(position >= goal_position and velocity >= goal_velocity)
    seam = seam or abs(position - goal_position) < 1e-6 or (position >= goal_position and abs(velocity - goal_velocity) < 1e-7)
    reward = 100.0 if terminated else 0.0
    reward -= math.pow(float(action), 2) * 0.1
    nxt = np.array([position, velocity])
    return nxt, nxt.copy(), reward, terminated, seam


def mountain_car_continuous_reset(rng):
    return np.array([rng.uniform(-0.6, -0.4), 0.0])


# ---- Pendulum-v1 (gymnasium/envs/classic_control/pendulum.py) -----------------------------------------------------
def pendulum_step(state, action):
    g, m, l, dt, max_speed, max_torque = 10.0, 1.0, 1.0, 0.05, 8.0, 2.0
    th, thdot = float(state[0]), float(state[1])
    u = float(np.clip(float(action), -max_torque, max_torque))
    angle_normalized = ((th + np.pi) % (2 * np.pi)) - np.pi
    costs = angle_normalized ** 2 + 0.1 * thdot ** 2 + 0.001 * (u ** 2)
    newthdot = thdot + (3 * g / (2 * l) * np.sin(th) + 3.0 / (m * l ** 2) * u) * dt
    seam = abs(abs(newthdot) - max_speed) < 1e-6
    newthdot = float(np.clip(newthdot, -max_speed, max_speed))
    newth = th + newthdot * dt
    nxt = np.array([newth, newthdot])
    obs = np.array([np.cos(newth), np.sin(newth), newthdot])
    return nxt, obs, -costs, False, seam


def pendulum_reset(rng):
    return np.array([rng.uniform(-np.pi, np.pi), rng.uniform(-1.0, 1.0)])


# ---- Acrobot-v1 (gymnasium/envs/classic_control/acrobot.py, book_or_nips = "book") --------------------------------
def _acrobot_dsdt(y, a):
    m1 = m2 = l1 = 1.0
    lc1 = lc2 = 0.5
    I1 = I2 = 1.0
    g = 9.8
    theta1, theta2, dtheta1, dtheta2 = y
    d1 = m1 * lc1 ** 2 + m2 * (l1 ** 2 + lc2 ** 2 + 2 * l1 * lc2 * math.cos(theta2)) + I1 + I2
    d2 = m2 * (lc2 ** 2 + l1 * lc2 * math.cos(theta2)) + I2
    phi2 = m2 * lc2 * g * math.cos(theta1 + theta2 - math.pi / 2.0)
    phi1 = (-m2 * l1 * lc2 * dtheta2 ** 2 * math.sin(theta2) - 2 * m2 * l1 * lc2 * dtheta2 * dtheta1 * math.sin(theta2)
            + (m1 * lc1 + m2 * l1) * g * math.cos(theta1 - math.pi / 2) + phi2)
    ddtheta2 = (a + d2 / d1 * phi1 - m2 * l1 * lc2 * dtheta1 ** 2 * math.sin(theta2) - phi2) / (m2 * lc2 ** 2 + I2 - d2 ** 2 / d1)
    ddtheta1 = -(d2 * ddtheta2 + phi1) / d1
    return np.array([dtheta1, dtheta2, ddtheta1, ddtheta2])


def _wrap(x, m, M):
    diff = M - m
    while x > M:
        x = x - diff
    while x < m:
        x = x + diff
    return x


def acrobot_step(state, action):
    dt = 0.2
    torque = float(action) - 1.0  # AVAIL_TORQUE = [-1, 0, +1]
    y0 = np.asarray(state, dtype=np.float64)
    k1 = _acrobot_dsdt(y0, torque)
    k2 = _acrobot_dsdt(y0 + dt / 2.0 * k1, torque)
    k3 = _acrobot_dsdt(y0 + dt / 2.0 * k2, torque)
    k4 = _acrobot_dsdt(y0 + dt * k3, torque)
    y = y0 + dt / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
    seam = min(abs(math.pi - abs(y[0])), abs(math.pi - abs(y[1]))) < 1e-5  # wrap decided either way: a 2 pi jump
    y[0] = _wrap(y[0], -math.pi, math.pi)
    y[1] = _wrap(y[1], -math.pi, math.pi)
    seam = seam or abs(abs(y[2]) - 4 * math.pi) < 1e-5 or abs(abs(y[3]) - 9 * math.pi) < 1e-5
    y[2] = min(max(y[2], -4 * math.pi), 4 * math.pi)
    y[3] = min(max(y[3], -9 * math.pi), 9 * math.pi)
    height = -math.cos(y[0]) - math.cos(y[1] + y[0])
    terminated = bool(height > 1.0)
    seam = seam or abs(height - 1.0) < 1e-4
    reward = 0.0 if terminated else -1.0
    obs = np.array([math.cos(y[0]), math.sin(y[0]), math.cos(y[1]), math.sin(y[1]), y[2], y[3]])
    return y, obs, reward, terminated, seam


def acrobot_reset(rng):
    return rng.uniform(-0.1, 0.1, size=4)


STEP = {MOUNTAIN_CAR: mountain_car_step, MOUNTAIN_CAR_CONTINUOUS: mountain_car_continuous_step, PENDULUM: pendulum_step,
        ACROBOT: acrobot_step}
RESET = {CARTPOLE: cartpole_reset, MOUNTAIN_CAR: mountain_car_reset, MOUNTAIN_CAR_CONTINUOUS: mountain_car_continuous_reset,
         PENDULUM: pendulum_reset, ACROBOT: acrobot_reset}


def policy(kind, rng, state, episode):
    """Actions that reach every branch: uniform random ones (beyond the Box bounds, to exercise the clamps) and, on
    every other episode, an energy-pumping rule that actually gets the car to the goal / the acrobot above the bar."""
    if kind == CARTPOLE:  # a PD rule that balances for the full 500 steps, or coin flips (falls within ~20 steps)
        if episode % 2 == 0:
            return int(state[2] + 0.5 * state[3] + 0.05 * state[0] + 0.1 * state[1] > 0)
        return int(rng.integers(0, 2))
    if kind == MOUNTAIN_CAR:  # push with the velocity (reaches the goal and, on the way, the left wall) or at random
        if episode % 2 == 0:
            return 2 if state[1] >= 0 else 0
        return int(rng.integers(0, 3))
    if kind == MOUNTAIN_CAR_CONTINUOUS:
        if episode % 2 == 0:
            return float(np.float32(math.copysign(rng.uniform(0.5, 1.2), state[1] if state[1] != 0 else 1.0)))
        return float(np.float32(rng.uniform(-1.2, 1.2)))
    if kind == PENDULUM:
        return float(np.float32(rng.uniform(-2.5, 2.5)))
    if episode % 2 == 0 and rng.random() < 0.9:
        return 2 if state[2] * math.cos(state[0]) < 0 else 0  # torque against the first link's swing (pumps energy)
    return int(rng.integers(0, 3))


def teacher_forced_trace(kind, steps, seed):
    """A `steps`-long trace in the reference's known-answer format: per step the f32 input state, the episode step
    count, the action, and the float64 answer (observation, reward, terminated, truncated, seam).  The NEXT input
    state is the answer rounded to f32 -- what an f32 implementation would hold -- or a fresh reset state after an
    episode ends."""
    rng = np.random.default_rng(seed)
    sd, od = STATE_DIM[kind], OBS_DIM[kind]
    out = {
        "state": np.zeros((steps, sd), np.float32), "count": np.zeros(steps, np.uint32),
        "action": np.zeros(steps, np.float32 if kind in (MOUNTAIN_CAR_CONTINUOUS, PENDULUM) else np.uint8),
        "next_state": np.zeros((steps, sd), np.float64), "obs": np.zeros((steps, od), np.float64),
        "reward": np.zeros(steps, np.float64), "terminated": np.zeros(steps, np.uint8),
        "truncated": np.zeros(steps, np.uint8), "seam": np.zeros(steps, np.uint8),
    }
    state = RESET[kind](rng).astype(np.float32)
    count, episode = 0, 0
    for t in range(steps):
        a = policy(kind, rng, state.astype(np.float64), episode)
        count_after = count + 1
        if kind == CARTPOLE:
            nxt, obs, reward, terminated, truncated, seam = cartpole_step(state.astype(np.float64), a, count)
        else:
            nxt, obs, reward, terminated, seam = STEP[kind](state.astype(np.float64), a)
            truncated = TIME_LIMIT[kind] > 0 and count_after >= TIME_LIMIT[kind]  # gymnasium.wrappers.TimeLimit
        # MountainCar-v0 has no limit: a random-policy episode is simply abandoned after 300 steps (a teacher-forced
        # trace may restart anywhere; the recorded answer of that step is untouched)
        abandon = kind == MOUNTAIN_CAR and count_after >= 300 and not terminated
        out["state"][t], out["count"][t], out["action"][t] = state, count, a
        out["next_state"][t], out["obs"][t], out["reward"][t] = nxt, obs, reward
        out["terminated"][t], out["truncated"][t], out["seam"][t] = terminated, truncated, seam
        if terminated or truncated or abandon:
            state, count, episode = RESET[kind](rng).astype(np.float32), 0, episode + 1
        else:
            state, count = nxt.astype(np.float32), count_after
    return out
