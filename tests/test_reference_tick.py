"""CPU: the restatements of the widening rows pinned against the reference's OWN controller.

oracle/_ref/libref_tick.so = the reference's ConvexMPCLocomotion, GaitGenerator, LegController, SwingLegController,
FootSwingTrajectory and DesiredCommand translation units (plus the three formulation files and qpOASES) compiled unchanged
against oracle/eigen_shim; oracle/ref_tick_probe.cpp ticks them the way FSMState_Walking::run does.  One simulated robot is
walked through more than a full gait cycle with a smoothly varying pose, and on every tick

  f-3  Gait::setIterations / mpc_gait          == scenarios.gait_phase / scenarios.mpc_gait
       run()'s touch-down heuristic            == the formula the closed-loop harness places feet with
  f-1  the `update_data_t` record the reference's updateMPCIfNeeded hands to solve_mpc
                                              == csrc/locomotion_host.cpp hmpc_prepare_record, BYTE FOR BYTE (live fields)
       the clamp write-back of world_position_desired                       == the host mirror's
  a16  f_ff = -rBody [F; M]                    == hmpc_wrench_to_feedforward, bit for bit
  f-2  J_force_moment and the joint torques    == oracle_leg_jacobian_fm / oracle_joint_torques
  f-4  swingLegController (both calls per tick) == oracle_swing_update: controller memory, touch-down point, pDes, vDes bit for
       bit; IK joint targets to 1e-12 rad
"""
import ctypes
import os

import numpy as np
import pytest

from conftest import GOLDEN
from hector_simulation_b200 import scenarios
from test_locomotion_host import Desired, LegData, StateEstimate, _host

DT, ITER_MPC, N = 0.001, 40, 10
DT_MPC = DT * ITER_MPC
N_TICKS = 520
COMMAND = dict(v_des=(0.3, 0.05), yaw_rate=0.2, roll=0.01, pitch=-0.02)
FIXTURE = os.path.join(GOLDEN, "ref_tick_walk.npz")   # the same ticks, recorded by tests/golden/make_ref_tick.py


@pytest.fixture(scope="module")
def ref(oracle):
    if not oracle.has_reference_tick():
        pytest.skip("oracle/_ref/libref_tick.so not built (needs /root/reference at build time)")
    return oracle


def _pose(k, rng_phase):
    """A smooth, deterministic pose sequence: forward drift, small body oscillations, moving joints."""
    t = k * DT
    a, b, c = rng_phase
    rpy = np.array([0.04 * np.sin(7 * t + a), 0.05 * np.sin(5 * t + b), 0.3 * t + 0.02 * np.sin(3 * t + c)])
    pos = np.array([0.25 * t + 0.01 * np.sin(9 * t), 0.02 * np.sin(4 * t + a), 0.55 + 0.01 * np.sin(6 * t + b)])
    vel = np.array([0.25 + 0.09 * np.cos(9 * t), 0.08 * np.cos(4 * t + a), 0.06 * np.cos(6 * t + b)])
    omega = np.array([0.28 * np.cos(7 * t + a), 0.25 * np.cos(5 * t + b), 0.3 + 0.06 * np.cos(3 * t + c)])
    raw = np.array([0.05 * np.sin(11 * t + i) for i in range(10)]) + np.tile([0.0, 0.02, 0.1, -0.2, 0.1], 2)
    return pos, rpy, vel, omega, raw.astype(np.float32)


def _state_record(o, pos, vel, quat, omega, cmd5, horizon=N):
    st = np.zeros((), dtype=scenarios.STATE_DTYPE)
    st["position"], st["vWorld"], st["orientation"], st["omegaWorld"] = pos, vel, quat, omega
    st["rpy"], st["leg_q"], st["leg_p"] = o["rpy"], o["leg_q"], o["leg_p"]
    st["state_des"] = cmd5
    st["world_position_desired"] = o["wpd_entry"][:2]
    st["gait"][: 2 * horizon] = o["mpc_table"]
    return st


# further cases (fixture ref_tick_cases.npz): the zero-command branches of the reference trajectory
# (ConvexMPCLocomotion.cpp:380-397 `== 0` tests) under the walking gait, and the standing gait
CASES = {
    "walk": dict(gait=2, n_ticks=N_TICKS, command=COMMAND, offsets=(0, 5), durations=(5, 5), pose=(0.3, 1.1, 2.0)),
    "walk_zero_command": dict(gait=2, n_ticks=200, command=dict(v_des=(0.0, 0.0), yaw_rate=0.0, roll=0.0, pitch=0.0),
                              offsets=(0, 5), durations=(5, 5), pose=(1.3, 0.2, 0.7)),
    # a command far from the motion: the set-point clamp (+-0.05 m, :338-346) engages within 100 ticks and the swing-leg
    # placement saturates at its FLOAT +-0.3 (fminf/fmaxf on doubles, SwingLegController.cpp:117-118)
    "walk_saturated": dict(gait=2, n_ticks=240, command=dict(v_des=(3.5, -3.2), yaw_rate=0.5, roll=0.0, pitch=0.03),
                           offsets=(0, 5), durations=(5, 5), pose=(0.9, 2.4, 1.6)),
    "stand": dict(gait=1, n_ticks=100, command=dict(v_des=(0.0, 0.0), yaw_rate=0.0, roll=0.0, pitch=0.0),
                  offsets=(0, 0), durations=(10, 10), pose=(2.1, 0.9, 0.1)),
}
CASES_FIXTURE = os.path.join(GOLDEN, "ref_tick_cases.npz")


def reference_ticks(O, case="walk"):
    """Tick the compiled reference controller through the pose sequence; yields its per-tick outputs."""
    c = CASES[case]
    ctl = O.ReferenceController(DT, ITER_MPC)
    for k in range(c["n_ticks"]):
        pos, rpy, vel, omega, raw = _pose(k, c["pose"])
        quat = scenarios.rpy_to_quat(rpy)
        yield ctl.run(c["gait"], pos, vel, quat, omega, raw, v_des_body=c["command"]["v_des"], yaw_rate=c["command"]["yaw_rate"],
                      roll=c["command"]["roll"], pitch=c["command"]["pitch"])
    ctl.close()


def committed_ticks(O, case="walk"):
    if case == "walk":
        return iter(np.load(FIXTURE)["ticks"].view(O.REFTICK_DTYPE).reshape(-1))
    return iter(np.load(CASES_FIXTURE)[case].view(O.REFTICK_DTYPE).reshape(-1))


@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("source", ["live", "committed"])
def test_walking_ticks_match_the_reference_controller(oracle, source, case):
    """live: libref_tick.so ticked here;  committed: the outputs it produced when the fixture was generated (runs on any
    machine, also without /root/reference)."""
    O = oracle
    C = CASES[case]
    if source == "live":
        if not O.has_reference_tick():
            pytest.skip("oracle/_ref/libref_tick.so not built (needs /root/reference at build time)")
        ticks = reference_ticks(O, case)
    else:
        ticks = committed_ticks(O, case)
    H = _host()
    swing = scenarios.make_swing(1)
    loop = np.zeros(1, dtype=scenarios.ROLLOUT_DTYPE)
    loop["gait_offset"], loop["gait_duration"] = C["offsets"], C["durations"]
    cmdd = C["command"]
    v_des, yaw_rate, roll, pitch = cmdd["v_des"], cmdd["yaw_rate"], cmdd["roll"], cmdd["pitch"]
    cmd5 = np.array([roll, pitch, v_des[0], v_des[1], yaw_rate])
    n_ticks, n_mpc, n_swing_checked, worst_ik = C["n_ticks"], 0, 0, 0.0
    n_clamped_setpoint = n_saturated_placement = 0
    g_stance, g_swing = C["durations"][0], N - C["durations"][0]          # Gait::_stance / _swing (GaitGenerator.cpp:13-14)
    for k, o in enumerate(ticks):
        pos, rpy, vel, omega, raw = _pose(k, C["pose"])
        quat = scenarios.rpy_to_quat(rpy)
        assert o["iteration_counter"] == k + 1

        # ---- f-3: gait ------------------------------------------------------------------------------------------
        it, ph = (k // ITER_MPC) % N, float(scenarios.gait_phase(k, ITER_MPC, N))
        assert o["gait_iteration"] == it and o["phase"] == ph
        assert np.array_equal(o["mpc_table"], scenarios.mpc_gait(N, C["offsets"], C["durations"], it).reshape(-1))

        # ---- the state estimate the probe wrote == what the restatements derive from the quaternion ---------------
        assert np.array_equal(o["rBody"].reshape(3, 3), _rbody_from_quat(quat))

        # ---- f-2: leg Jacobian ------------------------------------------------------------------------------------
        for leg in range(2):
            J = O.leg_jacobian_fm(o["leg_q"][5 * leg: 5 * leg + 5], leg)
            assert np.abs(J - o["J"][30 * leg: 30 * leg + 30].reshape(6, 5)).max() < 1e-15
            # the forward kinematics the scenario generator uses (LegController.cpp:190-194) on the same angles
            assert np.abs(scenarios.leg_fk(o["leg_q"][5 * leg: 5 * leg + 5], leg) - o["leg_p"][3 * leg: 3 * leg + 3]).max() < 1e-14

        # ---- f-1 / a16: on the ticks the reference solves ------------------------------------------------------------
        assert o["mpc_ran"] == (k % 5 == 0)
        if o["mpc_ran"]:
            n_mpc += 1
            st = _state_record(o, pos, vel, quat, omega, cmd5)
            rec = _host_record(H, st, o["rBody"])
            ref_rec = np.frombuffer(o["update_record"].tobytes(), dtype=scenarios.UPDATE_DTYPE)[0]
            for f in ("p", "v", "q", "w", "r", "joint_angles", "yaw", "weights", "Alpha_K"):
                assert rec[f].tobytes() == ref_rec[f].tobytes(), (k, f)
            assert rec["traj"][: 12 * N].tobytes() == ref_rec["traj"][: 12 * N].tobytes(), k
            assert np.array_equal(rec["gait"][: 2 * N], ref_rec["gait"][: 2 * N])
            # clamp write-back (ConvexMPCLocomotion.cpp:338-346)
            w = o["wpd_entry"].copy()
            for a in range(2):
                if w[a] - pos[a] > 0.05:
                    w[a] = pos[a] + 0.05
                if pos[a] - w[a] > 0.05:
                    w[a] = pos[a] - 0.05
            assert np.array_equal(w[:2], o["wpd"][:2])
            n_clamped_setpoint += int(not np.array_equal(w[:2], o["wpd_entry"][:2]))
            # f_ff = -rBody [F; M]
            ff = (ctypes.c_double * 12)()
            H.hmpc_wrench_to_feedforward(_dp(o["rBody"]), _dp(o["q_soln"][:12].copy()), ff)
            assert np.array_equal(np.array(ff[:]), o["f_ff"])
            # f-2: joint torques of the stance legs on this tick (J^T f_ff), the reference's own command path
            contact = np.array([[1 if np.any(o["ff_cmd"][6 * leg: 6 * leg + 6] != 0) else 0 for leg in range(2)]], np.int32)
            tau = O.joint_torques(o["q_soln"][:12], o["rBody"], o["leg_q"], contact)[0]
            assert np.allclose(tau, o["tau"], rtol=2e-6, atol=1e-5), (k, tau, o["tau"])   # lowCmd holds floats
        else:
            assert np.array_equal(o["wpd"], o["wpd_entry"])

        # ---- f-3: the touch-down heuristic of run() (ConvexMPCLocomotion.cpp:119-160) the closed-loop harness places feet
        #      with: hip projection + v * (remaining swing time) + clamp(0.5 v T_stance + 0.02 (v - v_des), +-0.4), z = 0.
        #      (run() never clears its firstSwing flags, so the remaining swing time stays dtMPC * _swing.)
        R = o["rBody"].reshape(3, 3).T
        vdw = R @ np.array([v_des[0], v_des[1], 0.0])
        for leg in range(2):
            rel = np.clip(vel[:2] * 0.5 * g_stance * DT_MPC + 0.02 * (vel[:2] - vdw[:2]), -0.4, 0.4).astype(np.float32)
            want = pos + R @ scenarios.hip_yaw_location(leg) + vel * (DT_MPC * g_swing)
            want[:2] += rel
            want[2] = -0.0
            assert np.abs(want - o["cmpc_pf"][3 * leg: 3 * leg + 3]).max() < 1e-7   # the reference clamps through floats (:152-153)

        # ---- f-4: swing-leg controller, called once per foot by run() (ConvexMPCLocomotion.cpp:218) ------------------
        st = _state_record(o, pos, vel, quat, omega, cmd5)
        states = np.array([st])
        for _ in range(2):
            cmd = O.swing_update(states, loop, np.array([o["phase"]]), swing, N, dt=DT, dt_swing=DT_MPC)
        assert np.array_equal(swing["swing_time"][0], o["swing_times"]), k
        assert np.array_equal(swing["first_swing"][0], o["first_swing"]), k
        assert np.array_equal(cmd["pf"][0], o["pf"]), k
        hipw = pos + o["rBody"].reshape(3, 3).T @ scenarios.hip_yaw_location(0) + vel * o["swing_times"][0]
        n_saturated_placement += int(abs(abs(o["pf"][0] - hipw[0]) - np.float32(0.3)) < 1e-12)
        for leg in range(2):
            sl = slice(3 * leg, 3 * leg + 3)
            if o["swing_states"][leg] > 0:
                n_swing_checked += 1
                assert cmd["swing"][0][leg] == 1
                assert np.array_equal(swing["p0"][0][sl], o["p0"][sl]), k
                assert np.array_equal(cmd["p_des"][0][sl], o["p_des"][sl]), k
                assert np.array_equal(cmd["v_des"][0][sl], o["v_des"][sl]), k
                d = np.abs(cmd["q_des"][0][5 * leg: 5 * leg + 5] - o["q_des"][5 * leg: 5 * leg + 5]).max()
                worst_ik = max(worst_ik, d)
            else:
                assert cmd["swing"][0][leg] == 0
    assert n_mpc == n_ticks // 5 and n_swing_checked > (300 if case == "walk" else (100 if g_swing else -1))
    if case == "walk_saturated":
        assert n_clamped_setpoint > 20 and n_saturated_placement > 50
    if not g_swing:
        assert n_swing_checked == 0
    assert worst_ik < 1e-12, worst_ik


def test_standing_tick_is_the_physically_sane_stand(ref):
    """configs[0] through the reference's whole controller: Fz = 47.84 N per foot."""
    b = scenarios.stand_inputs(N)
    ctl = ref.ReferenceController(DT, ITER_MPC)
    o = ctl.run(1, b["p"], b["v"], b["q"], b["w"], np.zeros(10, np.float32))
    ctl.close()
    assert o["mpc_ran"] == 1
    assert abs(o["q_soln"][2] - 47.84) < 0.05 and abs(o["q_soln"][5] - 47.84) < 0.05
    assert np.allclose(o["leg_q"], b["q_leg"], atol=1e-6) and np.allclose(o["leg_p"], b["leg_p"].reshape(-1), atol=1e-9)
    rec = np.frombuffer(o["update_record"].tobytes(), dtype=scenarios.UPDATE_DTYPE)[0]
    mine = scenarios.to_record(b, N)
    for f in ("p", "v", "q", "w", "joint_angles", "weights", "Alpha_K", "yaw"):
        assert np.array_equal(rec[f], mine[f]), f
    assert np.allclose(rec["r"], mine["r"], atol=1e-7) and np.allclose(rec["traj"][:120], mine["traj"][:120], atol=1e-6)


def test_reference_controller_links_against_the_product_library(oracle):
    """INTEGRATION.md's claim at link level: the reference's controller objects (ConvexMPCLocomotion.cpp & co., compiled
    unchanged) link with --no-undefined against libhector_mpc_b200.so in place of convexMPC_interface.cpp + SolverMPC.cpp +
    RobotState.cpp + qpOASES, importing exactly the boundary symbols of convexMPC_interface.h:39-43.  (Ticking it needs a
    GPU: tests/test_zz_device_vs_reference_vectors.py.)"""
    import shutil
    import subprocess

    if not oracle.has_reference_tick_dropin():
        pytest.skip("oracle/_ref/libref_tick_b200.so not built (needs /root/reference at build time)")
    L = oracle.tick_dropin_lib()           # loads libhector_mpc_b200.so through its RUNPATH
    for sym in ("reftick_create", "reftick_run", "reftick_destroy"):
        assert hasattr(L, sym)
    if shutil.which("nm"):
        und = subprocess.run(["nm", "-D", "--undefined-only", oracle._TICK_DROPIN_LIB_PATH], capture_output=True, text=True).stdout
        boundary = {l.split()[-1] for l in und.splitlines() if l.split()[-1] in
                    ("setup_problem", "update_problem_data", "get_solution", "update_solver_settings")}
        assert boundary == {"setup_problem", "update_problem_data", "get_solution"}   # what ConvexMPCLocomotion.cpp calls
        assert "solve_mpc" not in und and "qpOASES" not in und


# ---- helpers -----------------------------------------------------------------------------------------------------
def _dp(a):
    return np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _rbody_from_quat(q):
    """orientation_tools.h:182-200, the expression order of the reference (R, then transposed)."""
    e0, e1, e2, e3 = (float(x) for x in q)
    R = np.array([[1 - 2 * (e2 * e2 + e3 * e3), 2 * (e1 * e2 - e0 * e3), 2 * (e1 * e3 + e0 * e2)],
                  [2 * (e1 * e2 + e0 * e3), 1 - 2 * (e1 * e1 + e3 * e3), 2 * (e2 * e3 - e0 * e1)],
                  [2 * (e1 * e3 - e0 * e2), 2 * (e2 * e3 + e0 * e1), 1 - 2 * (e1 * e1 + e2 * e2)]])
    return R.T


def _host_record(H, st, rBody):
    se = StateEstimate()
    se.position[:] = st["position"]; se.orientation[:] = st["orientation"]; se.rpy[:] = st["rpy"]
    se.rBody[:] = np.asarray(rBody).reshape(-1)
    se.omegaWorld[:] = st["omegaWorld"]; se.vWorld[:] = st["vWorld"]
    legs = (LegData * 2)()
    for leg in range(2):
        legs[leg].q[:] = st["leg_q"][5 * leg: 5 * leg + 5]
        legs[leg].p[:] = st["leg_p"][3 * leg: 3 * leg + 3]
    cmd = Desired()
    cmd.stateDes[3], cmd.stateDes[4], cmd.stateDes[6], cmd.stateDes[7], cmd.stateDes[11] = st["state_des"]
    wpd = (ctypes.c_double * 2)(*st["world_position_desired"])
    table = st["gait"][: 2 * N].astype(np.int32)
    rec = np.zeros(1, dtype=scenarios.UPDATE_DTYPE)
    H.hmpc_prepare_record(ctypes.byref(se), legs, ctypes.byref(cmd), wpd, table.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                          ctypes.c_int(N), ctypes.c_double(DT_MPC), rec.ctypes.data_as(ctypes.c_void_p), None)
    return rec[0]


@pytest.mark.parametrize("case", list(CASES))
def test_feedforward_gate_follows_the_gait_sub_phases(oracle, case):
    """ConvexMPCLocomotion.cpp:199-266: a foot receives feedforwardForce = f_ff on the ticks where its swing sub-phase is 0
    and its contact sub-phase positive — from the continuous Gait::_phase of THIS tick, not from the contact table of the last
    MPC update.  The host mirror's Gait (getContactSubPhase / getSwingSubPhase restated) reproduces the command the
    reference's controller left in commands[].feedforwardForce on every recorded tick; gating on the MPC table instead would
    differ on the ticks around touch-down / lift-off."""
    O = oracle
    C = CASES[case]
    H = _host()
    off = (ctypes.c_int * 2)(*C["offsets"])
    dur = (ctypes.c_int * 2)(*C["durations"])
    prev = np.zeros(12)
    table_gate_differs = 0
    last_table = None
    for k, o in enumerate(committed_ticks(O, case)):
        cs, ss = (ctypes.c_double * 2)(), (ctypes.c_double * 2)()
        H.hloco_gait_subphases(N, off, dur, ITER_MPC, k, cs, ss)
        # the swing controller reads the same sub-phase (standing gait, tick 0: 0/0 = NaN in both, "not swinging")
        assert np.array_equal(np.array(ss[:]), o["swing_states"], equal_nan=True), k
        exp, exp_table = prev.copy(), prev.copy()
        if k % 5 == 0:
            last_table = o["mpc_table"].copy()
        for leg in range(2):
            sl = slice(6 * leg, 6 * leg + 6)
            if not ss[leg] > 0 and cs[leg] > 0:
                exp[sl] = o["f_ff"][sl]
            if last_table[leg] == 1:
                exp_table[sl] = o["f_ff"][sl]
        assert np.array_equal(exp, o["ff_cmd"]), k
        table_gate_differs += int(not np.array_equal(exp_table, o["ff_cmd"]))
        prev = o["ff_cmd"].copy()
    if case == "walk":
        assert table_gate_differs > 0      # the distinction is real on this sequence
