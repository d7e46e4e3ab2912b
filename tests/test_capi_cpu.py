"""CPU: the C-ABI shared library loads, exports every declared symbol, does its host-only byte work,
and fails LOUDLY without a GPU (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden
from hector_simulation_b200 import interface, scenarios


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "hector_mpc_b200.h")).read()
    return re.findall(r"HMPC_EXTERNC\s+[\w\s\*]+?\b(\w+)\s*\(", hdr)


def test_library_exports_every_declared_symbol():
    L = interface.lib()
    names = _declared_symbols()
    assert len(names) >= 14 and set(interface.EXPORTS) <= set(names)
    for n in names:
        assert hasattr(L, n), n
    # the four entry points of the reference boundary (convexMPC_interface.h:39-43)
    for n in ("setup_problem", "update_problem_data", "get_solution", "update_solver_settings"):
        assert n in names


def test_record_bytes_formula():
    # 216 + 98*N algorithmic bytes in+out per QP: input part is 216 + 50*N
    for N in (5, 10, 16):
        raw = (54 + 12 * N) * 4 + 2 * N
        assert raw == 216 + 50 * N
        assert interface.record_bytes(N) == (raw + 15) // 16 * 16
    assert interface.record_bytes(10) == 720
    assert interface.record_bytes(0) == 0 and interface.record_bytes(19) == 0


def test_pack_records_layout():
    g = load_golden("cfg3_h10")
    recs = g["records"][:5]
    N = 10
    packed = interface.pack_records(recs, N)
    assert packed.shape == (5, 720)
    for i in range(5):
        f = packed[i, : (54 + 12 * N) * 4].view(np.float32)
        r = recs[i]
        expect = np.concatenate([r["p"], r["v"], r["q"], r["w"], r["r"], r["joint_angles"], [r["yaw"]], r["weights"],
                                 r["Alpha_K"], r["traj"][: 12 * N]]).astype(np.float32)
        assert np.array_equal(f.view(np.uint32), expect.view(np.uint32))
        assert np.array_equal(packed[i, (54 + 12 * N) * 4: (54 + 12 * N) * 4 + 2 * N], r["gait"][: 2 * N])
        assert (packed[i, (54 + 12 * N) * 4 + 2 * N:] == 0).all()


def test_get_solution_is_zero_before_first_solve():
    # convexMPC_interface.cpp:107
    assert interface.get_solution(0) == 0.0 and interface.get_solution(119) == 0.0


def test_argument_errors_are_reported():
    L = interface.lib()
    assert L.hmpc_create(0, 10, 0) is None
    assert b"max_batch" in L.hmpc_last_error()
    assert L.hmpc_create(8, 17, 0) is None
    assert L.hmpc_pack_records(None, 1, 10, None) == interface.HMPC_ERR_ARG


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: this test is about the GPU-less box")
    with pytest.raises(interface.HmpcError) as e:
        interface.BatchedMPC(4, 10)
    assert "no usable CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_scenarios_double_to_float_narrowing():
    # update_problem_data narrows doubles to floats (convexMPC_interface.cpp:87-99)
    b = scenarios.stand_inputs(10)
    rec = scenarios.to_record(b, 10)
    assert rec["p"].dtype == np.float32 and np.array_equal(rec["p"], b["p"].astype(np.float32))
    assert np.array_equal(rec["traj"][:120], b["state_trajectory"].astype(np.float32))
    assert (rec["gait"][:20] == 1).all() and (rec["gait"][20:] == 0).all()
