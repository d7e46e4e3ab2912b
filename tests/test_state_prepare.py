"""Row f-1 on the device: hmpc_prepare_device / hmpc_solve_batch_states against the host mirror of the
reference's data preparation (csrc/locomotion_host.cpp hmpc_prepare_record = ConvexMPCLocomotion.cpp:283-406, itself
checked against the numpy restatement in test_locomotion_host.py)."""
import ctypes
import os

import numpy as np
import pytest

from conftest import ROOT
from hector_simulation_b200 import interface, scenarios
from test_locomotion_host import Desired, LegData, StateEstimate, _host


def _host_prepared(states, horizon, dt=0.04):
    """update_data_t records of `states` built by the host mirror of updateMPCIfNeeded."""
    L = _host()
    recs = np.zeros(len(states), dtype=scenarios.UPDATE_DTYPE)
    for i, st in enumerate(states):
        se = StateEstimate()
        se.position[:] = st["position"]; se.orientation[:] = st["orientation"]; se.rpy[:] = st["rpy"]
        se.rBody[:] = scenarios.quat_to_R(st["orientation"]).T.reshape(-1)
        se.omegaWorld[:] = st["omegaWorld"]; se.vWorld[:] = st["vWorld"]
        legs = (LegData * 2)()
        for leg in range(2):
            legs[leg].q[:] = st["leg_q"][5 * leg: 5 * leg + 5]
            legs[leg].p[:] = st["leg_p"][3 * leg: 3 * leg + 3]
        cmd = Desired()
        cmd.stateDes[3], cmd.stateDes[4], cmd.stateDes[6], cmd.stateDes[7], cmd.stateDes[11] = st["state_des"]
        wpd = (ctypes.c_double * 2)(*st["world_position_desired"])
        table = st["gait"][: 2 * horizon].astype(np.int32)
        L.hmpc_prepare_record(ctypes.byref(se), legs, ctypes.byref(cmd), wpd, table.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                              ctypes.c_int(horizon), ctypes.c_double(dt), recs[i:i + 1].ctypes.data_as(ctypes.c_void_p), None)
    return recs


def test_state_record_layout():
    assert scenarios.STATE_DTYPE.itemsize == 352
    assert scenarios.STATE_DTYPE.fields["gait"][1] == 39 * 8
    header = open(os.path.join(ROOT, "include", "hector_mpc_b200.h")).read()
    assert "struct hmpc_state_t" in header and "hmpc_solve_batch_states" in header and "hmpc_prepare_device" in header


def test_states_describe_the_same_tick_as_records():
    """to_state + host preparation reproduces make_batch's records (the numpy restatement) to float rounding."""
    recs, inputs = scenarios.make_batch(3, 16, horizon=10)
    states = scenarios.make_states(inputs, 10)
    host = _host_prepared(states, 10)
    for k in ("p", "v", "q", "w", "joint_angles", "weights", "Alpha_K", "gait", "yaw"):
        assert np.array_equal(host[k], recs[k]), k
    assert np.allclose(host["r"], recs["r"], atol=1e-7) and np.allclose(host["traj"], recs["traj"], atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("horizon,cfg,batch", [(10, 3, 96), (5, 4, 32), (16, 4, 32), (10, 1, 4)])
def test_device_preparation_is_bit_identical_to_host(horizon, cfg, batch):
    import torch

    _, inputs = scenarios.make_batch(cfg, batch, horizon=horizon)
    states = scenarios.make_states(inputs, horizon)
    host = interface.pack_records(_host_prepared(states, horizon), horizon)
    mpc = interface.BatchedMPC(batch, horizon)
    d_states = torch.from_numpy(states.view(np.uint8).reshape(batch, 352)).cuda()
    d_rec = torch.full((batch, host.shape[1]), 0xAB, dtype=torch.uint8, device="cuda")
    mpc.prepare_device(d_states, batch, d_rec)
    torch.cuda.synchronize()
    dev = d_rec.cpu().numpy()
    assert dev.shape == host.shape
    diff = np.nonzero(dev != host)
    words = sorted(set((diff[1] // 4).tolist()))
    assert diff[0].size == 0, f"{diff[0].size} differing bytes in {len(set(diff[0].tolist()))} robots; float words {words[:20]}"
    mpc.close()


@pytest.mark.gpu
def test_solve_batch_states_equals_record_path():
    horizon, batch = 10, 700  # two chunks
    _, inputs = scenarios.make_batch(3, batch, horizon=horizon, seed=77)
    states = scenarios.make_states(inputs, horizon)
    recs = _host_prepared(states, horizon)
    mpc = interface.BatchedMPC(batch, horizon)
    w_rec, tau_rec, s_rec = mpc.solve_batch_torques(recs)
    w_st, tau_st, s_st = mpc.solve_batch_states(states, torques=True)
    assert np.array_equal(s_rec, s_st)
    assert np.array_equal(w_rec, w_st) and np.array_equal(tau_rec, tau_st)
    assert (interface.status_code(s_st) == 0).all()
    # and a small batch (single chunk, no helper threads)
    w1, s1 = mpc.solve_batch_states(states[:5])
    assert np.array_equal(w1, w_rec[:5])
    mpc.close()
