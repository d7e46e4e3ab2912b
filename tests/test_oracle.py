"""CPU: the oracle against its own golden vectors and against first principles (no GPU)."""
import numpy as np
import pytest

from conftest import load_golden, rel_err
from hector_simulation_b200 import scenarios

CASES = ["cfg1_h10", "cfg2_h10", "cfg3_h10", "cfg4_h5", "cfg4_h16"]


def test_update_data_layout(oracle):
    # byte-compatible with convexMPC_interface.h:19-37
    assert scenarios.UPDATE_DTYPE.itemsize == 3016
    assert oracle.UPDATE_DTYPE == scenarios.UPDATE_DTYPE
    f = scenarios.UPDATE_DTYPE.fields
    assert f["traj"][1] == 4 * 42 and f["gait"][1] == 4 * 486 and f["max_iterations"][1] == 2980 and f["rho"][1] == 2984


@pytest.mark.parametrize("name", CASES)
def test_formulation_reproduces_golden_bitwise(oracle, name):
    g = load_golden(name)
    setup = oracle.make_setup(g["horizon"])
    for i in range(g["H"].shape[0]):
        f = oracle.formulate_f32(g["records"][i], setup)
        for k in ("H", "g", "Fblk", "lb", "ub"):
            assert np.array_equal(f[k].view(np.uint32), g[k][i].view(np.uint32)), (name, i, k)


@pytest.mark.parametrize("name", CASES)
def test_solver_reproduces_golden(oracle, name):
    if not oracle.has_qpoases():
        pytest.skip("oracle built without the reference's qpOASES")
    g = load_golden(name)
    q, info = oracle.solve_batch(g["records"], oracle.make_setup(g["horizon"]))
    assert np.array_equal(info, g["info"])
    assert np.array_equal(q, g["q_soln"])  # same binary, same inputs -> identical doubles


def test_stand_is_physically_sane():
    # SURVEY.md §8c self-check: symmetric double support, Fz0 = Fz1 = 47.84 N, My = -0.968 N m
    g = load_golden("cfg1_h10")
    u0 = g["q_soln"][0, :12]
    assert abs(u0[2] - 47.84) < 0.01 and abs(u0[5] - 47.84) < 0.01
    assert abs(u0[2] - u0[5]) < 1e-5
    assert abs(u0[7] + 0.9678) < 1e-3 and abs(u0[10] + 0.9678) < 1e-3
    assert g["info"][0, 2] == 120 and g["info"][0, 3] == 160


def test_swing_variables_are_exactly_zero():
    # SolverMPC.cpp:723-726: eliminated variables are written as 0
    g = load_golden("cfg2_h10")
    N = g["horizon"]
    q = g["q_soln"].reshape(-1, N, 12)
    gait = g["records"]["gait"][:, : 2 * N].reshape(-1, N, 2)
    for leg in range(2):
        cols = [3 * leg + c for c in range(3)] + [6 + 3 * leg + c for c in range(3)]
        assert (q[:, :, cols][gait[:, :, leg] == 0] == 0.0).all()
        assert (np.abs(q[:, :, 3 * leg + 2][gait[:, :, leg] == 1]) > 0).any()
    assert (g["info"][:, 2] == 60).all() and (g["info"][:, 3] == 80).all()


@pytest.mark.parametrize("name", ["cfg2_h10", "cfg3_h10", "cfg4_h5"])
def test_golden_solutions_satisfy_kkt(oracle, name):
    """The qpOASES optimum is a KKT point of the reduced QP it was given (fp64 check)."""
    g = load_golden(name)
    setup = oracle.make_setup(g["horizon"])
    for i in range(0, min(8, len(g["records"]))):
        Q = oracle.reduced_qp(g["records"][i], setup)
        x = g["q_soln"][i][Q["var_ind"]]
        Ax = Q["A"] @ x
        scale = max(1.0, np.abs(x).max())
        assert (Ax >= Q["lb"] - 1e-6 * scale).all() and (Ax <= Q["ub"] + 1e-6 * scale).all()
        Hs = np.triu(Q["H"]) + np.triu(Q["H"], 1).T
        grad = Hs @ x + Q["g"]
        act_lo = np.abs(Ax - Q["lb"]) < 1e-6 * scale
        act_hi = np.abs(Ax - Q["ub"]) < 1e-6 * scale
        Aact = Q["A"][act_lo | act_hi]
        # stationarity: grad in the range of the active rows, with the right multiplier signs
        lam, *_ = np.linalg.lstsq(Aact.T, grad, rcond=None) if len(Aact) else (np.zeros(0),)
        res = grad - (Aact.T @ lam if len(Aact) else 0)
        assert np.linalg.norm(res) <= 2e-5 * max(1.0, np.linalg.norm(grad), np.linalg.norm(Q["g"]))


def test_numpy_dual_active_set_matches_qpoases(oracle):
    """Independent solver (the kernel's algorithm in numpy fp64) agrees with the reference's qpOASES."""
    from oracle import qp_dual_active_set as G

    g = load_golden("cfg3_h10")
    setup = oracle.make_setup(10)
    for i in range(0, 12):
        Q = oracle.reduced_qp(g["records"][i], setup)
        x, inf = G.solve(Q["H"], Q["g"], Q["A"], Q["lb"], Q["ub"])
        assert inf["status"] == 0
        full = np.zeros(120)
        full[Q["var_ind"]] = x
        assert rel_err(full[None], g["q_soln"][i][None], 12)[0] < 2e-5
        assert inf["iters"] == g["info"][i, 1]  # same number of working-set changes as nWSR


def test_gait_tables():
    # Gait::mpc_gait, GaitGenerator.cpp:85-103 with walking(10,(0,5),(5,5)) / standing(10,(0,0),(10,10))
    t = scenarios.walking_table(10, 0).reshape(10, 2)
    assert (t[:5] == [1, 0]).all() and (t[5:] == [0, 1]).all()
    t3 = scenarios.walking_table(10, 3).reshape(10, 2)
    assert (t3[:2] == [1, 0]).all() and (t3[2:7] == [0, 1]).all() and (t3[7:] == [1, 0]).all()
    assert (scenarios.standing_table(10) == 1).all()


def test_stand_pose_feet_are_symmetric():
    b = scenarios.stand_inputs(10)
    r = b["r"]  # [x0,x1,y0,y1,z0,z1]
    assert abs(r[0] - r[1]) < 1e-12 and abs(r[2] + r[3]) < 1e-12 and abs(r[4] - r[5]) < 1e-12
    assert -0.56 < r[4] < -0.40  # feet below the CoM (spawn pose has bent knees)
