"""ctypes loader for the parity oracle — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this module.  It wraps oracle/_ref/liboracle_mpc.so: the Eigen-free restatement of the
reference's ``solve_mpc`` (hector_control/ConvexMPC/SolverMPC.cpp:371-732) linked against the
reference's own vendored qpOASES 3.2 built unchanged from /root/reference (oracle/Makefile).

Parity status: the restatement is pinned against the reference's own formulation sources compiled
unchanged against oracle/eigen_shim (oracle/_ref/libref_mpc.so, `ref_solve` below): in trig-as-compiled
mode it reproduces them bit for bit (tests/test_reference_compiled.py).  What stays a restatement is the
arithmetic of Eigen itself (absent here, unpinned by the reference) — see the header of
solve_mpc_oracle.cpp.  Solver half: the reference's own code.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "liboracle_mpc.so")

K_MAX_GAIT_SEGMENTS = 36

# numpy mirror of `update_data_t` (convexMPC_interface.h:19-37), C layout incl. padding.
UPDATE_DTYPE = np.dtype(
    [
        ("p", "<f4", 3),
        ("v", "<f4", 3),
        ("q", "<f4", 4),
        ("w", "<f4", 3),
        ("r", "<f4", 6),
        ("joint_angles", "<f4", 10),
        ("yaw", "<f4"),
        ("weights", "<f4", 12),
        ("traj", "<f4", 12 * K_MAX_GAIT_SEGMENTS),
        ("Alpha_K", "<f4", 12),
        ("gait", "u1", K_MAX_GAIT_SEGMENTS),
        ("hack_pad", "u1", 1000),
        ("max_iterations", "<i4"),
        ("rho", "<f8"),
        ("sigma", "<f8"),
        ("solver_alpha", "<f8"),
        ("terminate", "<f8"),
    ],
    align=True,
)
assert UPDATE_DTYPE.itemsize == 3016, UPDATE_DTYPE.itemsize

SETUP_DTYPE = np.dtype([("dt", "<f4"), ("mu", "<f4"), ("f_max", "<f4"), ("horizon", "<i4")], align=True)


def build(force: bool = False) -> str:
    """Run oracle/Makefile (compiles qpOASES from /root/reference when present)."""
    if force or not os.path.exists(_LIB_PATH) or os.path.isdir("/root/reference"):
        subprocess.run(["make", "-s", "-C", _HERE, "-j8"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        L.oracle_has_qpoases.restype = ctypes.c_int
        L.oracle_sizeof_update_data.restype = ctypes.c_size_t
        assert L.oracle_sizeof_update_data() == UPDATE_DTYPE.itemsize
        _lib = L
    return _lib


def has_qpoases() -> bool:
    return bool(lib().oracle_has_qpoases())


def make_setup(horizon: int = 10, dt: float = 0.04, mu: float = 0.25, f_max: float = 500.0) -> np.ndarray:
    s = np.zeros(1, dtype=SETUP_DTYPE)
    s["dt"], s["mu"], s["f_max"], s["horizon"] = dt, mu, f_max, horizon
    return s


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _mode(assembly_fp64: bool, trig_as_compiled: bool, gemv_by4: bool = False, gemm_kc: int = 0) -> ctypes.c_int:
    """Bit mask of solve_mpc_oracle.cpp: 1 = formulation in double, 2 = trig as the reference's TU resolves it,
    4 = matrix-vector products grouped like Eigen 3.3's gemv kernel (sensitivity probe)."""
    return ctypes.c_int(int(bool(assembly_fp64)) | (2 if trig_as_compiled else 0) | (4 if gemv_by4 else 0) | ((int(gemm_kc) & 0xff) << 8))


def solve_batch(records: np.ndarray, setup: np.ndarray, assembly_fp64: bool = False, trig_as_compiled: bool = False,
                gemv_by4: bool = False, gemm_kc: int = 0):
    """-> (q_soln [n,12N] f64, info [n,4] i32 = {rc, nWSR, nv_red, nc_red})."""
    records = np.ascontiguousarray(records, dtype=UPDATE_DTYPE)
    n = records.shape[0]
    N = int(setup["horizon"][0])
    q = np.zeros((n, 12 * N), dtype=np.float64)
    info = np.zeros((n, 4), dtype=np.int32)
    if not has_qpoases():
        raise RuntimeError("oracle was built without qpOASES (no /root/reference and no prebuilt oracle/_ref)")
    lib().oracle_solve_batch(_p(records), ctypes.c_int(n), _p(setup), _mode(assembly_fp64, trig_as_compiled, gemv_by4, gemm_kc), _p(q), _p(info))
    return q, info


def time_solves(records: np.ndarray, setup: np.ndarray, total: int) -> np.ndarray:
    """Per-solve wall seconds of `total` single-thread oracle solves (round-robin over records)."""
    records = np.ascontiguousarray(records, dtype=UPDATE_DTYPE)
    out = np.zeros(total, dtype=np.float64)
    lib().oracle_time_solves(_p(records), ctypes.c_int(records.shape[0]), _p(setup), ctypes.c_int(total), _p(out))
    return out


def formulate_f32(record: np.ndarray, setup: np.ndarray, trig_as_compiled: bool = False) -> dict:
    """Un-reduced fp32 QP data + intermediates of ONE record (the reference's arithmetic)."""
    record = np.ascontiguousarray(record, dtype=UPDATE_DTYPE).reshape(1)
    N = int(setup["horizon"][0])
    n = 12 * N
    out = dict(
        H=np.zeros((n, n), np.float32), g=np.zeros(n, np.float32), Fblk=np.zeros((16, 12), np.float32),
        lb=np.zeros(16 * N, np.float32), ub=np.zeros(16 * N, np.float32), x0=np.zeros(13, np.float32),
        Acd=np.zeros((13, 13), np.float32), Bcd=np.zeros((13, 12), np.float32), Rfoot=np.zeros((2, 3, 3), np.float32),
        R=np.zeros((3, 3), np.float32), A_qp=np.zeros((13 * N, 13), np.float32),
    )
    lib().oracle_formulate_f32_mode(_p(record), _p(setup), _mode(False, trig_as_compiled), _p(out["H"]), _p(out["g"]), _p(out["Fblk"]), _p(out["lb"]),
                                    _p(out["ub"]), _p(out["x0"]), _p(out["Acd"]), _p(out["Bcd"]), _p(out["Rfoot"]),
                                    _p(out["R"]), _p(out["A_qp"]))
    return out


def reduced_qp(record: np.ndarray, setup: np.ndarray, assembly_fp64: bool = False, trig_as_compiled: bool = False) -> dict:
    """The reduced QP exactly as handed to qpOASES (SolverMPC.cpp:644-697), doubles."""
    record = np.ascontiguousarray(record, dtype=UPDATE_DTYPE).reshape(1)
    N = int(setup["horizon"][0])
    n, m = 12 * N, 16 * N
    H = np.zeros(n * n); g = np.zeros(n); A = np.zeros(m * n); lb = np.zeros(m); ub = np.zeros(m)
    vi = np.zeros(n, np.int32); ci = np.zeros(m, np.int32); nc = ctypes.c_int(0)
    L = lib()
    L.oracle_reduced_qp.restype = ctypes.c_int
    nv = L.oracle_reduced_qp(_p(record), _p(setup), _mode(assembly_fp64, trig_as_compiled), _p(H), _p(g), _p(A), _p(lb), _p(ub),
                             _p(vi), _p(ci), ctypes.byref(nc))
    nc = nc.value
    return dict(H=H[: nv * nv].reshape(nv, nv).copy(), g=g[:nv].copy(), A=A[: nc * nv].reshape(nc, nv).copy(),
                lb=lb[:nc].copy(), ub=ub[:nc].copy(), var_ind=vi[:nv].copy(), con_ind=ci[:nc].copy())


def leg_jacobian_fm(q5, leg: int) -> np.ndarray:
    """J_force_moment (6x5) of LegController.cpp:130-166 for one leg (q5 = LegController's angles)."""
    q5 = np.ascontiguousarray(q5, dtype=np.float64)
    J = np.zeros((6, 5))
    lib().oracle_leg_jacobian_fm(_p(q5), ctypes.c_int(leg), _p(J))
    return J


def joint_torques(wrench12, rBody, q_leg, contact) -> np.ndarray:
    """tau [n,10] = J^T (-rBody [F;M]) per stance leg (LegController.cpp:57-63, ConvexMPCLocomotion.cpp:419-440)."""
    wrench12 = np.ascontiguousarray(wrench12, dtype=np.float64).reshape(-1, 12)
    n = wrench12.shape[0]
    rBody = np.ascontiguousarray(rBody, dtype=np.float64).reshape(n, 9)
    q_leg = np.ascontiguousarray(q_leg, dtype=np.float64).reshape(n, 10)
    contact = np.ascontiguousarray(contact, dtype=np.int32).reshape(n, 2)
    tau = np.zeros((n, 10))
    lib().oracle_joint_torques(_p(wrench12), _p(rBody), _p(q_leg), _p(contact), ctypes.c_int(n), _p(tau))
    return tau


def swing_update(states: np.ndarray, loop: np.ndarray, phase: np.ndarray, swing: np.ndarray, n_iterations: int,
                 dt: float = 0.001, dt_swing: float = 0.04) -> np.ndarray:
    """One swingLegController::updateSwingLeg for every robot (swing_leg_oracle.cpp); `swing` is updated in place.
    Arrays use the hmpc_state_t / hmpc_rollout_t / hmpc_swing_t layouts; returns hmpc_swing_cmd_t records."""
    n = states.shape[0]
    assert states.dtype.itemsize == 352 and loop.dtype.itemsize == 80 and swing.dtype.itemsize == 72
    assert states.flags.c_contiguous and loop.flags.c_contiguous and swing.flags.c_contiguous
    phase = np.ascontiguousarray(phase, dtype=np.float64)
    cmd = np.zeros(n, dtype=np.dtype([("pf", "<f8", 6), ("p_des", "<f8", 6), ("v_des", "<f8", 6), ("q_des", "<f8", 10), ("swing", "<i4", 2)]))
    assert cmd.dtype.itemsize == 232
    lib().oracle_swing_update(_p(states), _p(loop), _p(phase), _p(swing), ctypes.c_int(n), ctypes.c_int(n_iterations),
                              ctypes.c_double(dt), ctypes.c_double(dt_swing), _p(cmd))
    return cmd


# ---- the reference's own formulation sources, compiled against oracle/eigen_shim (oracle/Makefile) ------------
_REF_LIB_PATH = os.path.join(_HERE, "_ref", "libref_mpc.so")
_ref_lib = None


def has_reference_build() -> bool:
    """True when oracle/_ref/libref_mpc.so exists (built where /root/reference is present; it travels)."""
    return os.path.exists(_REF_LIB_PATH)


def ref_lib() -> ctypes.CDLL:
    global _ref_lib
    if _ref_lib is None:
        L = ctypes.CDLL(_REF_LIB_PATH)
        L.refshim_sizeof_update_data.restype = ctypes.c_size_t
        assert L.refshim_sizeof_update_data() == UPDATE_DTYPE.itemsize
        _ref_lib = L
    return _ref_lib


def ref_solve(records: np.ndarray, setup: np.ndarray, formulation: bool = False):
    """The reference's `resize_qp_mats` + `solve_mpc` (SolverMPC.cpp, compiled unchanged) on each record.
    The reference's c2qp loops are fixed at 10 (quirk Q1), so only horizon 10 is meaningful.
    -> q_soln [n,12N] f64, and with formulation=True a dict of the reference's file-scope QP matrices per record."""
    records = np.ascontiguousarray(records, dtype=UPDATE_DTYPE)
    n = records.shape[0]
    N = int(setup["horizon"][0])
    assert N == 10, "the reference formulation is hard-wired to horizon 10 (SolverMPC.cpp:148,161,180)"
    nv, nc = 12 * N, 16 * N
    q = np.zeros((n, nv))
    out = None
    if formulation:
        out = dict(H=np.zeros((n, nv, nv), np.float32), g=np.zeros((n, nv), np.float32), A=np.zeros((n, nc, nv), np.float32),
                   lb=np.zeros((n, nc), np.float32), ub=np.zeros((n, nc), np.float32), x0=np.zeros((n, 13), np.float32),
                   A_qp=np.zeros((n, 13 * N, 13), np.float32), A_ct=np.zeros((n, 13, 13), np.float32),
                   B_ct=np.zeros((n, 13, 12), np.float32))
    L = ref_lib()
    null = ctypes.c_void_p(0)
    for i in range(n):
        rec = records[i:i + 1]
        if formulation:
            L.refshim_solve(_p(rec), _p(setup), _p(q[i]), _p(out["H"][i]), _p(out["g"][i]), _p(out["A"][i]), _p(out["lb"][i]),
                            _p(out["ub"][i]), _p(out["x0"][i]), _p(out["A_qp"][i]), _p(out["A_ct"][i]), _p(out["B_ct"][i]))
        else:
            L.refshim_solve(_p(rec), _p(setup), _p(q[i]), null, null, null, null, null, null, null, null, null)
    return (q, out) if formulation else q


def ref_boundary_solve(b: dict, horizon: int = 10, dt: float = 0.04, mu: float = 0.25, f_max: float = 500.0) -> np.ndarray:
    """`setup_problem` + `update_problem_data` + `get_solution` of the compiled reference on one set of boundary
    inputs (doubles, the dict layout of scenarios.boundary_inputs)."""
    d = lambda k: np.ascontiguousarray(b[k], dtype=np.float64)
    gait = np.ascontiguousarray(b["gait"], dtype=np.int32)
    q = np.zeros(12 * horizon)
    arrs = [d(k) for k in ("p", "v", "q", "w", "r", "joint_angles")]
    w8, traj, alpha = d("weights"), d("state_trajectory"), d("Alpha_K")
    ref_lib().refshim_boundary_solve(ctypes.c_double(dt), ctypes.c_int(horizon), ctypes.c_double(mu), ctypes.c_double(f_max),
                                     *[_p(a) for a in arrs], ctypes.c_double(float(b["yaw"])), _p(w8), _p(traj), _p(alpha),
                                     _p(gait), _p(q))
    return q


# ---- the reference's controller tick (rows f-1..f-4), compiled against oracle/eigen_shim (oracle/Makefile) ----------
_TICK_LIB_PATH = os.path.join(_HERE, "_ref", "libref_tick.so")
_tick_lib = None

# numpy mirror of `reftick_out_t` (oracle/ref_tick_probe.cpp)
REFTICK_DTYPE = np.dtype(
    [
        ("rBody", "<f8", 9), ("rpy", "<f8", 3), ("leg_q", "<f8", 10), ("leg_p", "<f8", 6), ("J", "<f8", 60),
        ("wpd_before", "<f8", 3), ("wpd_entry", "<f8", 3), ("wpd", "<f8", 3), ("phase", "<f8"), ("gait_iteration", "<i4"),
        ("mpc_table", "<i4", 20), ("mpc_ran", "<i4"), ("iteration_counter", "<i4"), ("q_soln", "<f8", 120),
        ("f_ff", "<f8", 12), ("swing_states", "<f8", 2), ("swing_times", "<f8", 2), ("first_swing", "<i4", 2),
        ("p0", "<f8", 6), ("pf", "<f8", 6), ("q_des", "<f8", 10), ("p_des", "<f8", 6), ("v_des", "<f8", 6),
        ("ff_cmd", "<f8", 12), ("cmpc_pf", "<f8", 6), ("tau", "<f8", 10), ("update_record", "u1", 3016),
    ],
    align=True,
)


_TICK_DROPIN_LIB_PATH = os.path.join(_HERE, "_ref", "libref_tick_b200.so")
_tick_dropin_lib = None


def has_reference_tick() -> bool:
    return os.path.exists(_TICK_LIB_PATH)


def has_reference_tick_dropin() -> bool:
    """oracle/_ref/libref_tick_b200.so: the reference's controller objects linked against the product library."""
    return os.path.exists(_TICK_DROPIN_LIB_PATH)


def tick_dropin_lib() -> ctypes.CDLL:
    global _tick_dropin_lib
    if _tick_dropin_lib is None:
        L = ctypes.CDLL(_TICK_DROPIN_LIB_PATH)
        L.reftick_sizeof_out.restype = ctypes.c_size_t
        L.reftick_create.restype = ctypes.c_void_p
        assert L.reftick_sizeof_out() == REFTICK_DTYPE.itemsize
        _tick_dropin_lib = L
    return _tick_dropin_lib


def tick_lib() -> ctypes.CDLL:
    global _tick_lib
    if _tick_lib is None:
        L = ctypes.CDLL(_TICK_LIB_PATH)
        L.reftick_sizeof_out.restype = ctypes.c_size_t
        L.reftick_create.restype = ctypes.c_void_p
        assert L.reftick_sizeof_out() == REFTICK_DTYPE.itemsize, (L.reftick_sizeof_out(), REFTICK_DTYPE.itemsize)
        _tick_lib = L
    return _tick_lib


class ReferenceController:
    """One robot's walking controller of the reference (ConvexMPCLocomotion + LegController + swing-leg controller +
    gait), ticked the way FSMState_Walking::run does.  State persists across ticks like the reference's objects;
    the MPC itself is process-global in the reference (one controller solving at a time)."""

    def __init__(self, dt: float = 0.001, iterations_between_mpc: int = 40, drop_in: bool = False):
        """drop_in=True: the same controller objects with libhector_mpc_b200.so behind setup_problem /
        update_problem_data / get_solution (needs a GPU to run; `update_record` stays zero)."""
        self._L = tick_dropin_lib() if drop_in else tick_lib()
        self._h = ctypes.c_void_p(self._L.reftick_create(ctypes.c_double(dt), ctypes.c_int(iterations_between_mpc)))
        assert self._h.value, "reftick_create failed"

    def run(self, gait_number, position, vWorld, orientation, omegaWorld, motor_q, motor_dq=None, v_des_body=(0.0, 0.0),
            yaw_rate=0.0, roll=0.0, pitch=0.0) -> np.ndarray:
        d = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        position, vWorld, orientation, omegaWorld, vdb = d(position), d(vWorld), d(orientation), d(omegaWorld), d(v_des_body)
        mq = np.ascontiguousarray(motor_q, dtype=np.float32)
        mdq = np.zeros(10, np.float32) if motor_dq is None else np.ascontiguousarray(motor_dq, dtype=np.float32)
        out = np.zeros(1, dtype=REFTICK_DTYPE)
        self._L.reftick_run(self._h, ctypes.c_int(gait_number), _p(position), _p(vWorld), _p(orientation), _p(omegaWorld),
                            _p(mq), _p(mdq), _p(vdb), ctypes.c_double(yaw_rate), ctypes.c_double(roll), ctypes.c_double(pitch),
                            _p(out))
        return out[0]

    def close(self):
        if self._h is not None and self._h.value:
            self._L.reftick_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
