"""GPU (-m gpu): the device kernels of the widening rows against the COMMITTED outputs of the reference's own controller
(tests/golden/ref_tick_walk.npz: ConvexMPCLocomotion / LegController / SwingLegController / GaitGenerator compiled
unchanged, 520 walking ticks — see tests/test_reference_tick.py for how they were produced and for the CPU-side checks
of the host restatements against the same vectors).  Nothing here needs oracle/_ref or /root/reference.

  f-1  hmpc_prepare_device   == the `update_data_t` records the reference's updateMPCIfNeeded built, byte for byte
  f-4  hmpc_swing_device     == the reference's swing-leg controller, tick by tick (two calls per tick, as run() makes them)
  f-2  hmpc_solve_batch_ex   == the reference's first-step wrench (1e-4 contract) and the joint torques it commanded
"""
import numpy as np
import pytest

from conftest import rel_err
from hector_simulation_b200 import interface, scenarios
from test_reference_tick import COMMAND, DT, DT_MPC, FIXTURE, ITER_MPC, N, _pose, _state_record

pytestmark = pytest.mark.gpu


def _ticks():
    from oracle import oracle_py  # only for the record layout of the fixture (no oracle code runs)

    return np.load(FIXTURE)["ticks"].view(oracle_py.REFTICK_DTYPE).reshape(-1)


def _cmd5():
    return np.array([COMMAND["roll"], COMMAND["pitch"], COMMAND["v_des"][0], COMMAND["v_des"][1], COMMAND["yaw_rate"]])


def _states(ticks, which):
    out = np.zeros(len(which), dtype=scenarios.STATE_DTYPE)
    for n, k in enumerate(which):
        pos, rpy, vel, omega, _ = _pose(int(k), (0.3, 1.1, 2.0))
        out[n] = _state_record(ticks[k], pos, vel, scenarios.rpy_to_quat(rpy), omega, _cmd5())
    return out


def _ref_records(ticks, which):
    return np.array([np.frombuffer(ticks[k]["update_record"].tobytes(), dtype=scenarios.UPDATE_DTYPE)[0] for k in which])


def test_device_preparation_reproduces_the_reference_controllers_records():
    import torch

    ticks = _ticks()
    which = np.nonzero(ticks["mpc_ran"])[0]
    states = _states(ticks, which)
    want = interface.pack_records(_ref_records(ticks, which), N)
    B = len(which)
    mpc = interface.BatchedMPC(B, N)
    d_states = torch.from_numpy(states.view(np.uint8).reshape(B, 352)).cuda()
    d_rec = torch.full((B, want.shape[1]), 0xAB, dtype=torch.uint8, device="cuda")
    mpc.prepare_device(d_states, B, d_rec, dt_mpc=DT_MPC)
    torch.cuda.synchronize()
    got = d_rec.cpu().numpy()
    mpc.close()
    diff = np.nonzero(got != want)
    assert diff[0].size == 0, f"{diff[0].size} differing bytes; float words {sorted(set((diff[1] // 4).tolist()))[:20]}"


def test_device_swing_controller_follows_the_reference_controller():
    import torch

    ticks = _ticks()
    mpc = interface.BatchedMPC(1, N)
    loop = np.zeros(1, dtype=scenarios.ROLLOUT_DTYPE)
    loop["gait_offset"], loop["gait_duration"] = (0, 5), (5, 5)
    d_loop = torch.from_numpy(loop.view(np.uint8).reshape(1, 80)).cuda()
    d_sw = torch.from_numpy(scenarios.make_swing(1).view(np.uint8).reshape(1, 72)).cuda()
    d_cmd = torch.zeros((1, 232), dtype=torch.uint8, device="cuda")
    worst = dict(pf=0.0, p0=0.0, p_des=0.0, v_des=0.0, q_des=0.0, swing_time=0.0)
    checked = 0
    for k in range(len(ticks)):
        o = ticks[k]
        st = _states(ticks, [k])
        d_st = torch.from_numpy(st.view(np.uint8).reshape(1, 352)).cuda()
        d_ph = torch.tensor([float(o["phase"])], dtype=torch.float64, device="cuda")
        for _ in range(2):   # ConvexMPCLocomotion::run calls updateSwingLeg once per foot (ConvexMPCLocomotion.cpp:218)
            mpc.swing_device(d_st, d_loop, d_ph, d_sw, 1, d_cmd, dt=DT, dt_swing=DT_MPC)
        torch.cuda.synchronize()
        cmd = d_cmd.cpu().numpy().view(scenarios.SWING_CMD_DTYPE).reshape(1)[0]
        sw = d_sw.cpu().numpy().view(scenarios.SWING_DTYPE).reshape(1)[0]
        assert np.array_equal(sw["first_swing"], o["first_swing"]), k
        worst["swing_time"] = max(worst["swing_time"], float(np.abs(sw["swing_time"] - o["swing_times"]).max()))
        worst["pf"] = max(worst["pf"], float(np.abs(cmd["pf"] - o["pf"]).max()))
        for leg in range(2):
            in_swing = o["swing_states"][leg] > 0
            assert int(cmd["swing"][leg]) == int(in_swing), (k, leg)
            if in_swing:
                checked += 1
                s3, s5 = slice(3 * leg, 3 * leg + 3), slice(5 * leg, 5 * leg + 5)
                worst["p0"] = max(worst["p0"], float(np.abs(sw["p0"][s3] - o["p0"][s3]).max()))
                worst["p_des"] = max(worst["p_des"], float(np.abs(cmd["p_des"][s3] - o["p_des"][s3]).max()))
                worst["v_des"] = max(worst["v_des"], float(np.abs(cmd["v_des"][s3] - o["v_des"][s3]).max()))
                worst["q_des"] = max(worst["q_des"], float(np.abs(cmd["q_des"][s5] - o["q_des"][s5]).max()))
    mpc.close()
    print("swing controller, device vs the reference's compiled controller, worst abs diff:", worst, "legs checked:", checked)
    assert checked > 300
    assert worst["swing_time"] < 1e-12 and worst["p0"] < 1e-12 and worst["pf"] < 1e-12
    assert worst["p_des"] < 1e-12 and worst["v_des"] < 1e-12
    assert worst["q_des"] < 1e-9   # asin/acos: last bits differ between libm and the CUDA math library


def test_device_solve_and_torques_follow_the_reference_controller():
    ticks = _ticks()
    which = np.nonzero(ticks["mpc_ran"])[0]
    recs = _ref_records(ticks, which)
    mpc = interface.BatchedMPC(len(which), N)
    w, tau, st = mpc.solve_batch_torques(recs)
    mpc.close()
    assert (interface.status_code(st) == 0).all()
    q_ref = np.array([ticks[k]["q_soln"] for k in which])
    assert rel_err(w, q_ref, 12).max() < 1e-4                    # the contract, against the reference's own solve_mpc
    assert (w[q_ref == 0.0] == 0.0).all()
    tau_ref = np.array([ticks[k]["tau"] for k in which]).reshape(-1, 2, 5)
    commanded = np.array([[np.any(ticks[k]["ff_cmd"][6 * leg: 6 * leg + 6] != 0) for leg in range(2)] for k in which])
    got = tau.reshape(-1, 2, 5)[commanded]
    want = tau_ref[commanded]
    assert commanded.sum() > 90
    # the reference's torques pass through float motor commands; ours start from the float-narrowed record
    err = np.linalg.norm(got - want, axis=1) / np.maximum(np.linalg.norm(want, axis=1), 1.0)
    assert err.max() < 5e-4, err.max()


def test_reference_controller_runs_on_the_gpu_library():
    """The drop-in, end to end: the reference's OWN controller objects (ConvexMPCLocomotion, LegController, swing-leg
    controller, gait — compiled unchanged) linked against libhector_mpc_b200.so instead of their MPC files, ticked through
    the walking sequence, against what the same controller produced with the reference's own solve_mpc + qpOASES."""
    from oracle import oracle_py as O   # loader of the test-side library only; no oracle arithmetic runs

    if not O.has_reference_tick_dropin():
        pytest.skip("oracle/_ref/libref_tick_b200.so not built (needs /root/reference at build time)")
    from test_reference_tick import CASES

    ticks = _ticks()
    c = CASES["walk"]
    ctl = O.ReferenceController(DT, ITER_MPC, drop_in=True)
    worst = dict(wrench=0.0, f_ff=0.0, tau=0.0)
    n_mpc = 0
    for k in range(len(ticks)):
        want = ticks[k]
        pos, rpy, vel, omega, raw = _pose(k, c["pose"])
        o = ctl.run(c["gait"], pos, vel, scenarios.rpy_to_quat(rpy), omega, raw, v_des_body=c["command"]["v_des"],
                    yaw_rate=c["command"]["yaw_rate"], roll=c["command"]["roll"], pitch=c["command"]["pitch"])
        # everything around the solver is the same machine code: identical
        for f in ("leg_q", "leg_p", "J", "wpd", "phase", "mpc_table", "swing_states", "swing_times", "first_swing", "p0", "pf",
                  "q_des", "cmpc_pf", "iteration_counter", "mpc_ran"):
            assert np.array_equal(o[f], want[f]), (k, f)
        if want["mpc_ran"]:
            n_mpc += 1
            scale = np.linalg.norm(want["q_soln"][:12])
            worst["wrench"] = max(worst["wrench"], float(np.linalg.norm(o["q_soln"][:12] - want["q_soln"][:12]) / scale))
            worst["f_ff"] = max(worst["f_ff"], float(np.linalg.norm(o["f_ff"] - want["f_ff"]) / np.linalg.norm(want["f_ff"])))
        nt = max(np.linalg.norm(want["tau"]), 1.0)
        worst["tau"] = max(worst["tau"], float(np.linalg.norm(o["tau"] - want["tau"]) / nt))
    ctl.close()
    print("reference controller on the GPU library vs on its own solver, worst relative gaps:", worst, "MPC ticks:", n_mpc)
    assert n_mpc == len(ticks) // 5
    assert worst["wrench"] < 1e-4 and worst["f_ff"] < 1e-4 and worst["tau"] < 1e-4
