"""Row f-4: the batched swing-leg controller (hmpc_swing_device) against the CPU restatement of
swingLegController::updateSwingLeg (oracle/swing_leg_oracle.cpp), plus property checks of the restatement itself."""
import numpy as np
import pytest

from hector_simulation_b200 import interface, scenarios

N = 10
IT_PER_MPC = 40  # iterationsBetweenMPC (FSMState_Walking.cpp:5)


def _robots(n, seed=None):
    _, inputs = scenarios.make_batch(5, n, horizon=N, seed=seed)
    return scenarios.make_rollout(inputs, N)


def test_swing_record_layouts():
    assert scenarios.SWING_DTYPE.itemsize == 72 and scenarios.SWING_DTYPE.fields["first_swing"][1] == 64
    assert scenarios.SWING_CMD_DTYPE.itemsize == 232 and scenarios.SWING_CMD_DTYPE.fields["swing"][1] == 224
    assert scenarios.SWING_CMD_DTYPE.fields["q_des"][1] == 144


def test_oracle_swing_trajectory_properties(oracle):
    """Bezier swing from the lift-off point to the planned touch-down, apex 0.15 m at mid swing, commands only for
    the leg in swing, swing-time countdown and firstSwing toggling as in SwingLegController.cpp:82-93,134-141."""
    states, loop = _robots(8)
    sw = scenarios.make_swing(8)
    R = [scenarios.quat_to_R(q) for q in states["orientation"]]
    lift = {}
    seen_mid = 0
    for it in range(0, 2 * IT_PER_MPC * N):  # two gait cycles at 1 kHz, robot states frozen
        ph = scenarios.gait_phase(np.full(8, it), IT_PER_MPC, N)
        before = sw.copy()
        cmd = oracle.swing_update(states, loop, ph, sw, N)
        for i in range(8):
            for leg in range(2):
                # walking(10, (0,5), (5,5)): leg 0 swings in phase (0.5, 1) (and at exactly 0: sub-phase 1), leg 1 in (0, 0.5]
                ph_i = ph[i]
                expect = (ph_i > 0.5 or ph_i == 0.0) if leg == 0 else (0.0 < ph_i <= 0.5)
                assert bool(cmd["swing"][i, leg]) == expect, (it, i, leg, ph_i)
                if not expect:
                    assert not cmd["q_des"][i, 5 * leg: 5 * leg + 5].any() and not cmd["p_des"][i, 3 * leg: 3 * leg + 3].any()
                    continue
                sub = (ph_i - 0.5) / 0.5 if leg == 0 and ph_i > 0.5 else (1.0 if leg == 0 else ph_i / 0.5)
                if before["first_swing"][i, leg]:
                    lift[(i, leg)] = sw["p0"][i, 3 * leg: 3 * leg + 3].copy()   # lift-off point latched on the first swing tick
                    assert sw["first_swing"][i, leg] == 0 and lift[(i, leg)][2] == 0.0
                p0, pf = lift[(i, leg)], cmd["pf"][i, 3 * leg: 3 * leg + 3]
                side = 1.0 if leg == 1 else -1.0
                p_world = states["position"][i] + R[i] @ (cmd["p_des"][i, 3 * leg: 3 * leg + 3] - np.array([-0.015, side * -0.055, 0.0]))
                b = sub ** 3 + 3 * sub ** 2 * (1 - sub)
                assert np.allclose(p_world[:2], p0[:2] + b * (pf[:2] - p0[:2]), atol=1e-12)
                assert -1e-12 <= p_world[2] <= 0.15 + 1e-12
                if abs(sub - 0.5) < 1e-12:
                    assert abs(p_world[2] - 0.15) < 1e-12
                    seen_mid += 1
                assert np.isfinite(cmd["q_des"][i]).all() and cmd["q_des"][i, 5 * leg] == 0.0
                assert np.allclose(cmd["v_des"][i, 3 * leg: 3 * leg + 3], -(R[i].T @ states["vWorld"][i]), atol=1e-14)
    assert seen_mid > 0
    # countdown: swingTimes restart at dtSwing * _swing = 0.2 s and fall by 1 ms per tick
    assert (sw["swing_time"] <= 0.2 + 1e-12).all() and (sw["swing_time"] > -1e-3 - 1e-12).all()


def test_oracle_touchdown_placement_formula(oracle):
    """Pf = position + rBody^T hip + vWorld * swingTime + clamp_float(1.75 * 0.5 v T_stance + 0.1 (v - v_des)), z = 0."""
    states, loop = _robots(16, seed=3)
    sw = scenarios.make_swing(16)
    cmd = oracle.swing_update(states, loop, scenarios.gait_phase(np.full(16, 123), IT_PER_MPC, N), sw, N)
    for i in range(16):
        R = scenarios.quat_to_R(states["orientation"][i])
        vdw = R @ np.array([states["state_des"][i, 2], states["state_des"][i, 3], 0.0])
        for leg in range(2):
            pf = states["position"][i] + R @ scenarios.hip_yaw_location(leg) + states["vWorld"][i] * sw["swing_time"][i, leg]
            for a in range(2):
                rel = 1.75 * states["vWorld"][i, a] * 0.5 * 5 * 0.04 + 0.1 * (states["vWorld"][i, a] - vdw[a])
                pf[a] += float(np.float32(min(max(np.float32(rel), np.float32(-0.3)), np.float32(0.3))))
            pf[2] = 0.0
            assert np.allclose(cmd["pf"][i, 3 * leg: 3 * leg + 3], pf, atol=1e-12)


def _to_dev(a):
    import torch

    return torch.from_numpy(a.view(np.uint8).reshape(a.shape[0], -1).copy()).cuda()


@pytest.mark.gpu
def test_device_swing_controller_matches_oracle(oracle):
    import torch

    B = 256
    mpc = interface.BatchedMPC(B, N)
    rng = np.random.default_rng(17)
    states, loop = _robots(B, seed=5)
    loop["gait_offset"][::7] = (0, 0)      # a few standing robots: durations = N, never in swing
    loop["gait_duration"][::7] = (N, N)
    sw_cpu = scenarios.make_swing(B)
    d_sw = _to_dev(sw_cpu)
    d_loop = _to_dev(loop)
    d_cmd = torch.zeros((B, 232), dtype=torch.uint8, device="cuda")
    counters = rng.integers(0, IT_PER_MPC * N, B)
    worst = dict(pf=0.0, p_des=0.0, v_des=0.0, q_des=0.0)
    for tick in range(80):
        if tick % 10 == 0 and tick:  # new robot states now and then (the controller memory carries over)
            st2, _ = _robots(B, seed=100 + tick)
            states[:] = st2
        counters = counters + rng.integers(1, 9, B)  # irregular strides hit boundaries and mid-phases
        phase = scenarios.gait_phase(counters, IT_PER_MPC, N)
        ref = oracle.swing_update(states, loop, phase, sw_cpu, N)
        mpc.swing_device(_to_dev(states), d_loop, torch.from_numpy(phase).cuda(), d_sw, B, d_cmd)
        torch.cuda.synchronize()
        got = d_cmd.cpu().numpy().view(scenarios.SWING_CMD_DTYPE).reshape(B)
        sw_dev = d_sw.cpu().numpy().view(scenarios.SWING_DTYPE).reshape(B)
        assert np.array_equal(got["swing"], ref["swing"]) and np.array_equal(sw_dev["first_swing"], sw_cpu["first_swing"])
        assert np.abs(sw_dev["swing_time"] - sw_cpu["swing_time"]).max() < 1e-14 and np.abs(sw_dev["p0"] - sw_cpu["p0"]).max() < 1e-13
        for k in worst:
            worst[k] = max(worst[k], float(np.abs(got[k] - ref[k]).max()))
    print("swing controller, device vs oracle, worst abs diff:", worst)
    assert worst["pf"] < 1e-12 and worst["p_des"] < 1e-12 and worst["v_des"] < 1e-12
    assert worst["q_des"] < 1e-9   # asin/acos: last-bit differences between libm and the CUDA math library
    mpc.close()
