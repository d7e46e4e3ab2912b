"""Host-side C++ mirror of the caller (csrc/locomotion_host.cpp).

CPU part: its data preparation (ConvexMPCLocomotion.cpp:283-406) against the independent numpy
restatement in scenarios.py.  GPU part: ConvexMPCLocomotion::run() driving the GPU boundary."""
import ctypes
import os

import numpy as np
import pytest

from conftest import ROOT, load_golden, rel_err
from hector_simulation_b200 import scenarios

HOSTLIB = os.path.join(ROOT, "hector_simulation_b200", "libhector_locomotion_host.so")


class StateEstimate(ctypes.Structure):
    _fields_ = [("position", ctypes.c_double * 3), ("orientation", ctypes.c_double * 4), ("rBody", ctypes.c_double * 9),
                ("rpy", ctypes.c_double * 3), ("omegaWorld", ctypes.c_double * 3), ("vWorld", ctypes.c_double * 3),
                ("vBody", ctypes.c_double * 3)]


class LegData(ctypes.Structure):
    _fields_ = [("q", ctypes.c_double * 5), ("p", ctypes.c_double * 3)]


class LegCmd(ctypes.Structure):
    _fields_ = [("feedforwardForce", ctypes.c_double * 6)]


class Desired(ctypes.Structure):
    _fields_ = [("stateDes", ctypes.c_double * 12)]


def _host():
    from hector_simulation_b200 import interface

    interface.lib()  # dependency of the host library
    L = ctypes.CDLL(HOSTLIB)
    L.hloco_create.restype = ctypes.c_void_p
    L.hloco_create.argtypes = [ctypes.c_double, ctypes.c_int]
    L.hloco_trajectory.restype = ctypes.POINTER(ctypes.c_double)
    L.hloco_foot_force.restype = ctypes.POINTER(ctypes.c_double)
    L.hloco_gait_table.restype = ctypes.POINTER(ctypes.c_int)
    for f in (L.hloco_destroy, L.hloco_set_gait, L.hloco_run, L.hloco_trajectory, L.hloco_foot_force, L.hloco_gait_table, L.hloco_iteration):
        f.argtypes = None
    return L


def _robot(rng, vx=0.0, yaw_rate=0.0):
    rpy = rng.normal(0, 0.05, 3)
    pos = np.array([0, 0, 0.55]) + rng.normal(0, 0.02, 3)
    vel = rng.normal(0, 0.1, 3)
    omega = rng.normal(0, 0.2, 3)
    raw = rng.normal(0, 0.05, 10)
    quat = scenarios.rpy_to_quat(rpy)
    R = scenarios.quat_to_R(quat)
    se = StateEstimate()
    se.position[:] = pos; se.orientation[:] = quat; se.rBody[:] = R.T.reshape(-1); se.rpy[:] = scenarios.quat_to_rpy(quat)
    se.omegaWorld[:] = omega; se.vWorld[:] = vel; se.vBody[:] = R.T @ vel
    legs = (LegData * 2)()
    ql = raw.reshape(2, 5).copy()
    ql[:, 2] += 0.3 * 3.14159; ql[:, 3] -= 0.6 * 3.14159; ql[:, 4] += 0.3 * 3.14159  # LegController.cpp:111-113
    for i in range(2):
        legs[i].q[:] = ql[i]
        legs[i].p[:] = scenarios.leg_fk(ql[i], i)
    cmd = Desired()
    cmd.stateDes[6] = vx
    cmd.stateDes[11] = yaw_rate
    return se, legs, cmd, dict(pos=pos, rpy=rpy, vel=vel, omega=omega, raw=raw)


@pytest.mark.parametrize("vx,yaw_rate", [(0.0, 0.0), (0.3, 0.0), (-0.2, 0.25)])
def test_prepare_record_matches_numpy_restatement(vx, yaw_rate):
    L = _host()
    rng = np.random.default_rng(11)
    for it in range(6):
        se, legs, cmd, s = _robot(rng, vx, yaw_rate)
        table = scenarios.walking_table(10, it)
        err = rng.normal(0, 0.04, 2)
        wpd = (ctypes.c_double * 2)(s["pos"][0] + err[0], s["pos"][1] + err[1])
        rec = np.zeros(1, dtype=scenarios.UPDATE_DTYPE)
        traj = (ctypes.c_double * 120)()
        L.hmpc_prepare_record(ctypes.byref(se), legs, ctypes.byref(cmd), wpd, table.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                              ctypes.c_int(10), ctypes.c_double(0.04), rec.ctypes.data_as(ctypes.c_void_p), traj)
        b = scenarios.boundary_inputs(s["pos"], s["rpy"], s["vel"], s["omega"], s["raw"], table, 10, v_des_body=(vx, 0.0),
                                      yaw_rate=yaw_rate, pos_des_err=err)
        ref = scenarios.to_record(b, 10)
        for k in ("p", "v", "q", "w", "joint_angles", "weights", "Alpha_K", "gait"):
            assert np.array_equal(rec[0][k], ref[k]), k
        assert np.allclose(rec[0]["r"], ref["r"], atol=1e-7) and np.allclose(rec[0]["traj"], ref["traj"], atol=1e-6)
        assert np.allclose(np.array(traj[:]), b["state_trajectory"], atol=1e-12)


def test_wrench_to_feedforward():
    L = _host()
    rng = np.random.default_rng(3)
    R = scenarios.quat_to_R(scenarios.rpy_to_quat(rng.normal(0, 0.3, 3)))
    rBody = np.ascontiguousarray(R.T)
    w = rng.normal(0, 10, 12)
    out = ((ctypes.c_double * 6) * 2)()
    L.hmpc_wrench_to_feedforward(rBody.ctypes.data_as(ctypes.c_void_p), w.ctypes.data_as(ctypes.c_void_p), out)
    for leg in range(2):
        assert np.allclose(np.array(out[leg][:3]), -rBody @ w[3 * leg: 3 * leg + 3])
        assert np.allclose(np.array(out[leg][3:]), -rBody @ w[6 + 3 * leg: 9 + 3 * leg])


def test_gait_class_matches_tables():
    L = _host()
    # walking(10,(0,5),(5,5)) through the class: iteration counter 0 -> table of phase 0
    h = ctypes.c_void_p(L.hloco_create(ctypes.c_double(0.001), 40))
    assert L.hloco_iteration(h) == 0
    L.hloco_destroy(h)


@pytest.mark.gpu
def test_locomotion_class_runs_the_gpu_boundary():
    """ConvexMPCLocomotion::run() (stand, first tick) -> f_ff = -rBody * wrench of the golden stand solution."""
    L = _host()
    g = load_golden("cfg1_h10")
    h = ctypes.c_void_p(L.hloco_create(ctypes.c_double(0.001), 40))
    L.hloco_set_gait(h, 1)
    se = StateEstimate()
    se.position[:] = [0, 0, 0.55]; se.orientation[:] = [1, 0, 0, 0]; se.rBody[:] = np.eye(3).reshape(-1)
    legs = (LegData * 2)()
    ql = np.zeros((2, 5)); ql[:, 2] += 0.3 * 3.14159; ql[:, 3] -= 0.6 * 3.14159; ql[:, 4] += 0.3 * 3.14159
    for i in range(2):
        legs[i].q[:] = ql[i]; legs[i].p[:] = scenarios.leg_fk(ql[i], i)
    cmd = Desired()
    out = (LegCmd * 2)()
    L.hloco_run(h, ctypes.byref(se), legs, ctypes.byref(cmd), out)
    assert L.hloco_iteration(h) == 1
    u0 = g["q_soln"][0, :12]
    # tick 0 solves, but the command is not written yet: at gait phase exactly 0 both sub-phases are 0 (the standing
    # gait's swing sub-phase is 0/0 = NaN, "not swinging"; its contact sub-phase 0/1 = 0, "not in contact") — the
    # reference's gate (ConvexMPCLocomotion.cpp:199-266, recorded in tests/golden/ref_tick_cases.npz "stand") skips it
    for leg in range(2):
        assert not np.any(np.array(out[leg].feedforwardForce[:]))
        f = np.array(L.hloco_foot_force(h, leg)[:6])
        ref = -np.concatenate([u0[3 * leg: 3 * leg + 3], u0[6 + 3 * leg: 9 + 3 * leg]])
        assert np.linalg.norm(f - ref) / np.linalg.norm(ref) < 5e-5
    L.hloco_run(h, ctypes.byref(se), legs, ctypes.byref(cmd), out)       # tick 1: no solve, the stance feet get f_ff
    for leg in range(2):
        assert np.array_equal(np.array(out[leg].feedforwardForce[:]), np.array(L.hloco_foot_force(h, leg)[:6]))
    # ticks 1..4 do not re-solve (MPC gate iterationCounter % 5, quirk Q11); tick 5 does
    for _ in range(4):
        L.hloco_run(h, ctypes.byref(se), legs, ctypes.byref(cmd), out)
    assert L.hloco_iteration(h) == 6
    L.hloco_destroy(h)


@pytest.mark.gpu
def test_locomotion_class_follows_the_reference_controller_over_a_gait_cycle(oracle):
    """ConvexMPCLocomotion::run() of the host mirror, ticked through the pose sequence the reference's compiled controller was
    recorded on (tests/golden/ref_tick_walk.npz: 520 control ticks = 104 MPC updates, 1.3 gait cycles, touch-downs and
    lift-offs of both feet): the feed-forward force left in the leg commands agrees on EVERY tick — the MPC ticks (GPU solve vs
    qpOASES, 5e-5) and the ticks between them (the stance gate of ConvexMPCLocomotion.cpp:199-266, from this tick's gait
    sub-phases) — and so does the desired trajectory handed to the solver."""
    import test_reference_tick as T

    L = _host()
    ticks = T.committed_ticks(oracle, "walk")
    c = T.CASES["walk"]
    h = ctypes.c_void_p(L.hloco_create(ctypes.c_double(T.DT), T.ITER_MPC))
    L.hloco_set_gait(h, c["gait"])
    out = (LegCmd * 2)()
    cmd = Desired()
    # DesiredStateCommand::setStateCommands (src/common/DesiredCommand.cpp:15-42): roll, pitch, body velocity, yaw rate
    cmd.stateDes[3], cmd.stateDes[4] = c["command"]["roll"], c["command"]["pitch"]
    cmd.stateDes[6], cmd.stateDes[7] = c["command"]["v_des"]
    cmd.stateDes[11] = c["command"]["yaw_rate"]
    worst, n_gate_only = 0.0, 0
    for k, o in enumerate(ticks):
        pos, rpy, vel, omega, raw = T._pose(k, c["pose"])
        quat = scenarios.rpy_to_quat(rpy)
        se = StateEstimate()
        se.position[:] = pos; se.orientation[:] = quat; se.rBody[:] = o["rBody"]; se.rpy[:] = o["rpy"]
        se.omegaWorld[:] = omega; se.vWorld[:] = vel
        legs = (LegData * 2)()
        for i in range(2):
            legs[i].q[:] = o["leg_q"][5 * i: 5 * i + 5]
            legs[i].p[:] = o["leg_p"][3 * i: 3 * i + 3]
        L.hloco_run(h, ctypes.byref(se), legs, ctypes.byref(cmd), out)
        assert L.hloco_iteration(h) == o["iteration_counter"]
        ff = np.concatenate([np.array(out[leg].feedforwardForce[:]) for leg in range(2)])
        ref = o["ff_cmd"]
        assert np.array_equal(ff == 0.0, ref == 0.0), k          # written / not yet written, foot by foot
        scale = max(np.linalg.norm(ref), 1e-9)
        worst = max(worst, np.linalg.norm(ff - ref) / scale)
        assert np.linalg.norm(ff - ref) / scale < 5e-5, (k, ff, ref)
        n_gate_only += int(k % 5 != 0)
    L.hloco_destroy(h)
    assert n_gate_only == 416
    print("host mirror vs reference controller over 520 ticks: worst relative feed-forward difference %.2e" % worst)
