"""Python face of libhector_mpc_b200.so (ctypes over the C-ABI in include/hector_mpc_b200.h).

Mirrors the reference's boundary, hector_control/ConvexMPC/convexMPC_interface.h:39-43 — same names,
argument order and semantics — and adds the batched calls.  All numerical work happens in the CUDA
library; this module only marshals pointers.  There is no CPU fallback: if the shared library or a
B200 is missing, calls raise.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from .scenarios import UPDATE_DTYPE

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhector_mpc_b200.so")

HMPC_OK, HMPC_ERR_ARG, HMPC_ERR_CUDA, HMPC_ERR_NOT_CONVERGED = 0, 1, 2, 3
ST_OK, ST_ITER_CAP, ST_WS_CAP, ST_INFEASIBLE, ST_NOT_SPD = 0, 1, 2, 3, 4

EXPORTS = [
    "setup_problem", "get_solution", "update_solver_settings", "update_problem_data",
    "hmpc_record_bytes", "hmpc_pack_records", "hmpc_create", "hmpc_destroy", "hmpc_last_error",
    "hmpc_set_problem", "hmpc_solve_batch", "hmpc_solve_device", "hmpc_launches_per_solve",
    "hmpc_assemble_device", "hmpc_class_config", "hmpc_solve_batch_ex", "hmpc_solve_device_ex",
    "hmpc_prepare_device", "hmpc_solve_batch_states", "hmpc_rollout_device", "hmpc_reset_warm_start",
    "hmpc_shard_unique_id", "hmpc_shard_init", "hmpc_solve_batch_sharded", "hmpc_shard_wait",
    "hmpc_pin_host_buffer", "hmpc_unpin_host_buffer", "hmpc_swing_device",
    "hmpc_reference_last_status", "hmpc_reference_last_rc",
]

SETUP_DTYPE = np.dtype([("dt", "<f4"), ("mu", "<f4"), ("f_max", "<f4"), ("horizon", "<i4")], align=True)

_lib = None


class HmpcError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Load the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HmpcError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = ctypes.CDLL(LIB_PATH)
        L.setup_problem.argtypes = [ctypes.c_double, ctypes.c_int, ctypes.c_double, ctypes.c_double]
        L.setup_problem.restype = None
        L.get_solution.argtypes = [ctypes.c_int]
        L.get_solution.restype = ctypes.c_double
        L.update_solver_settings.argtypes = [ctypes.c_int] + [ctypes.c_double] * 5
        L.update_solver_settings.restype = None
        dp = ctypes.POINTER(ctypes.c_double)
        L.update_problem_data.argtypes = [dp, dp, dp, dp, dp, dp, ctypes.c_double, dp, dp, dp, ctypes.POINTER(ctypes.c_int)]
        L.update_problem_data.restype = None
        L.hmpc_record_bytes.argtypes = [ctypes.c_int]
        L.hmpc_record_bytes.restype = ctypes.c_size_t
        L.hmpc_pack_records.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.hmpc_pack_records.restype = ctypes.c_int
        L.hmpc_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.hmpc_create.restype = ctypes.c_void_p
        L.hmpc_destroy.argtypes = [ctypes.c_void_p]
        L.hmpc_destroy.restype = None
        L.hmpc_last_error.restype = ctypes.c_char_p
        L.hmpc_set_problem.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.hmpc_set_problem.restype = ctypes.c_int
        L.hmpc_solve_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.hmpc_solve_batch.restype = ctypes.c_int
        L.hmpc_solve_device.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.hmpc_solve_device.restype = ctypes.c_int
        L.hmpc_launches_per_solve.argtypes = [ctypes.c_void_p]
        L.hmpc_launches_per_solve.restype = ctypes.c_int
        L.hmpc_assemble_device.argtypes = [ctypes.c_void_p] * 2 + [ctypes.c_int] + [ctypes.c_void_p] * 6
        L.hmpc_assemble_device.restype = ctypes.c_int
        L.hmpc_reference_last_status.restype = ctypes.c_int
        L.hmpc_reference_last_rc.restype = ctypes.c_int
        L.hmpc_solve_batch_ex.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 3
        L.hmpc_solve_batch_ex.restype = ctypes.c_int
        L.hmpc_solve_device_ex.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 4
        L.hmpc_solve_device_ex.restype = ctypes.c_int
        L.hmpc_prepare_device.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]
        L.hmpc_prepare_device.restype = ctypes.c_int
        L.hmpc_solve_batch_states.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_double] + [ctypes.c_void_p] * 3
        L.hmpc_solve_batch_states.restype = ctypes.c_int
        L.hmpc_rollout_device.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.hmpc_rollout_device.restype = ctypes.c_int
        L.hmpc_reset_warm_start.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.hmpc_reset_warm_start.restype = ctypes.c_int
        L.hmpc_shard_unique_id.argtypes = [ctypes.c_void_p]
        L.hmpc_shard_unique_id.restype = ctypes.c_int
        L.hmpc_shard_init.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.hmpc_shard_init.restype = ctypes.c_int
        L.hmpc_solve_batch_sharded.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.hmpc_solve_batch_sharded.restype = ctypes.c_int
        L.hmpc_shard_wait.argtypes = [ctypes.c_void_p]
        L.hmpc_shard_wait.restype = ctypes.c_int
        L.hmpc_pin_host_buffer.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
        L.hmpc_pin_host_buffer.restype = ctypes.c_int
        L.hmpc_unpin_host_buffer.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.hmpc_unpin_host_buffer.restype = ctypes.c_int
        L.hmpc_swing_device.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]
        L.hmpc_swing_device.restype = ctypes.c_int
        L.hmpc_class_config.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.hmpc_class_config.restype = ctypes.c_int
        _lib = L
    return _lib


def last_error() -> str:
    return lib().hmpc_last_error().decode()


def _check(rc: int, allow_not_converged: bool = False) -> int:
    if rc == HMPC_OK or (allow_not_converged and rc == HMPC_ERR_NOT_CONVERGED):
        return rc
    raise HmpcError(f"libhector_mpc_b200 rc={rc}: {last_error()}")


# ---------------------------------------------------------------------------------------------------
# Part 1: the reference's own entry points (convexMPC_interface.h:39-43)
# ---------------------------------------------------------------------------------------------------
def setup_problem(dt: float, horizon: int, mu: float, f_max: float) -> None:
    lib().setup_problem(dt, horizon, mu, f_max)


def update_problem_data(p, v, q, w, r, joint_angles, yaw, weights, state_trajectory, Alpha_K, gait) -> None:
    """Blocks until the solve is done (convexMPC_interface.cpp:83-103)."""
    def d(a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    keep = [d(a) for a in (p, v, q, w, r, joint_angles, weights, state_trajectory, Alpha_K)]
    g = np.ascontiguousarray(gait, dtype=np.int32)
    ptr = [k[1] for k in keep]
    lib().update_problem_data(ptr[0], ptr[1], ptr[2], ptr[3], ptr[4], ptr[5], float(yaw), ptr[6], ptr[7], ptr[8],
                              g.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))


def get_solution(index: int) -> float:
    return lib().get_solution(index)


def update_solver_settings(max_iter, rho, sigma, solver_alpha, terminate, use_jcqp) -> None:
    lib().update_solver_settings(max_iter, rho, sigma, solver_alpha, terminate, use_jcqp)


def reference_last_status() -> int:
    return lib().hmpc_reference_last_status()


def reference_last_rc() -> int:
    """Result of the last update_problem_data: 0, HMPC_ERR_NOT_CONVERGED, or the error the tick failed with."""
    return lib().hmpc_reference_last_rc()


# ---------------------------------------------------------------------------------------------------
# Part 2: batched interface
# ---------------------------------------------------------------------------------------------------
def record_bytes(horizon: int) -> int:
    return int(lib().hmpc_record_bytes(horizon))


def pack_records(records: np.ndarray, horizon: int) -> np.ndarray:
    """reference records -> packed device layout (uint8 [B, stride]); pure byte shuffling in C."""
    records = np.ascontiguousarray(records, dtype=UPDATE_DTYPE)
    out = np.zeros((records.shape[0], record_bytes(horizon)), dtype=np.uint8)
    _check(lib().hmpc_pack_records(records.ctypes.data, records.shape[0], horizon, out.ctypes.data))
    return out


def unpack_records(packed: np.ndarray, horizon: int) -> np.ndarray:
    """Inverse of pack_records (numpy): packed device records -> `update_data_t` records, e.g. to hand what a device
    loop logged to another consumer of the reference's record format."""
    packed = np.ascontiguousarray(packed, dtype=np.uint8).reshape(-1, record_bytes(horizon))
    n = packed.shape[0]
    f = packed[:, : (54 + 12 * horizon) * 4].copy().view(np.float32)
    out = np.zeros(n, dtype=UPDATE_DTYPE)
    out["p"], out["v"], out["q"], out["w"], out["r"] = f[:, 0:3], f[:, 3:6], f[:, 6:10], f[:, 10:13], f[:, 13:19]
    out["joint_angles"], out["yaw"], out["weights"], out["Alpha_K"] = f[:, 19:29], f[:, 29], f[:, 30:42], f[:, 42:54]
    out["traj"][:, : 12 * horizon] = f[:, 54: 54 + 12 * horizon]
    g0 = (54 + 12 * horizon) * 4
    out["gait"][:, : 2 * horizon] = packed[:, g0: g0 + 2 * horizon]
    return out


def status_code(s):
    return np.asarray(s) & 0xFF


def status_iters(s):
    return (np.asarray(s) >> 8) & 0xFFF


def status_nactive(s):
    return (np.asarray(s) >> 20) & 0xFF


def page_aligned(shape, dtype) -> np.ndarray:
    """A zeroed array that owns whole memory pages (start aligned, size rounded up): what a control loop should hand to
    BatchedMPC.pin / hmpc_pin_host_buffer — a C caller uses posix_memalign the same way."""
    page = os.sysconf("SC_PAGESIZE") if hasattr(os, "sysconf") else 4096
    dt = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dt.itemsize
    raw = np.zeros((nbytes + page - 1) // page * page + page, dtype=np.uint8)
    off = (-raw.ctypes.data) % page
    return raw[off:off + nbytes].view(dt).reshape(shape)


class BatchedMPC:
    """Context for `max_batch` robots of `horizon` steps on one GPU (hmpc_create / hmpc_destroy)."""

    def __init__(self, max_batch: int, horizon: int = 10, device: int = 0, dt: float = 0.04, mu: float = 0.25,
                 f_max: float = 500.0):
        self._h = lib().hmpc_create(max_batch, horizon, device)
        if not self._h:
            raise HmpcError(f"hmpc_create failed: {last_error()}")
        self.max_batch, self.horizon, self.device = max_batch, horizon, device
        self.set_problem(dt, mu, f_max)

    def set_problem(self, dt: float, mu: float, f_max: float) -> None:
        s = np.zeros(1, dtype=SETUP_DTYPE)
        s["dt"], s["mu"], s["f_max"], s["horizon"] = dt, mu, f_max, self.horizon
        _check(lib().hmpc_set_problem(self._h, s.ctypes.data))

    @property
    def launches_per_solve(self) -> int:
        return lib().hmpc_launches_per_solve(self._h)

    def class_config(self, cls: int) -> dict:
        out = np.zeros(6, dtype=np.int32)
        _check(lib().hmpc_class_config(self._h, cls, out.ctypes.data))
        return dict(zip(("threads", "smem_bytes", "qmax", "grid_cap", "nb_cap", "strip"), (int(v) for v in out)))

    def pin(self, *arrays: np.ndarray) -> None:
        """Register caller-owned arrays (records, wrench, status) for the in-place mode of solve_batch: the GPU then
        reads the records where they lie and writes the results where the caller wants them (hmpc_pin_host_buffer).
        The arrays must stay alive until unpin()/close().  Registration pins whole pages: allocate the arrays with
        page_aligned() so that no unrelated heap object shares their pages (a later cudaMemcpy of such a neighbour, partly
        inside a registered page range, fails with cudaErrorInvalidValue)."""
        for a in arrays:
            assert a.flags.c_contiguous
            _check(lib().hmpc_pin_host_buffer(self._h, a.ctypes.data, a.nbytes))

    def unpin(self, *arrays: np.ndarray) -> None:
        for a in arrays:
            _check(lib().hmpc_unpin_host_buffer(self._h, a.ctypes.data))

    def solve_batch(self, records: np.ndarray, strict: bool = True, out=None):
        """Host-buffer path: H2D + kernels + D2H inside.  -> (wrench [B,12N] f64, status [B] i32).
        `out=(wrench, status)` reuses caller-owned result arrays (what a C caller in a control loop does)."""
        if records.dtype != UPDATE_DTYPE or not records.flags.c_contiguous:
            records = np.ascontiguousarray(records, dtype=UPDATE_DTYPE)
        B = records.shape[0]
        if out is not None:
            wrench, status = out
            assert wrench.dtype == np.float64 and wrench.shape == (B, 12 * self.horizon) and wrench.flags.c_contiguous
            assert status.dtype == np.int32 and status.shape == (B,)
        else:
            wrench = np.zeros((B, 12 * self.horizon), dtype=np.float64)
            status = np.zeros(B, dtype=np.int32)
        _check(lib().hmpc_solve_batch(self._h, records.ctypes.data, B, wrench.ctypes.data, status.ctypes.data),
               allow_not_converged=not strict)
        return wrench, status

    def solve_batch_torques(self, records: np.ndarray, strict: bool = True):
        """Host path with the leg-controller epilogue: -> (wrench [B,12N], tau [B,10], status [B])."""
        records = np.ascontiguousarray(records, dtype=UPDATE_DTYPE)
        B = records.shape[0]
        wrench = np.zeros((B, 12 * self.horizon), dtype=np.float64)
        tau = np.zeros((B, 10), dtype=np.float64)
        status = np.zeros(B, dtype=np.int32)
        _check(lib().hmpc_solve_batch_ex(self._h, records.ctypes.data, B, wrench.ctypes.data, tau.ctypes.data, status.ctypes.data),
               allow_not_converged=not strict)
        return wrench, tau, status

    def solve_batch_states(self, states: np.ndarray, strict: bool = True, torques: bool = False, out=None, dt_mpc: float = 0.04):
        """Row f-1: `hmpc_state_t` records in, data preparation on the device.  -> (wrench, [tau,] status)."""
        from .scenarios import STATE_DTYPE

        if states.dtype != STATE_DTYPE or not states.flags.c_contiguous:
            states = np.ascontiguousarray(states, dtype=STATE_DTYPE)
        B = states.shape[0]
        if out is not None:
            wrench, status = out
            assert wrench.dtype == np.float64 and wrench.shape == (B, 12 * self.horizon) and wrench.flags.c_contiguous
            assert status.dtype == np.int32 and status.shape == (B,)
        else:
            wrench = np.zeros((B, 12 * self.horizon), dtype=np.float64)
            status = np.zeros(B, dtype=np.int32)
        tau = np.zeros((B, 10), dtype=np.float64) if torques else None
        _check(lib().hmpc_solve_batch_states(self._h, states.ctypes.data, B, dt_mpc, wrench.ctypes.data,
                                             tau.ctypes.data if torques else None, status.ctypes.data),
               allow_not_converged=not strict)
        return (wrench, tau, status) if torques else (wrench, status)

    def prepare_device(self, d_states, B: int, d_records, stream=None, dt_mpc: float = 0.04) -> None:
        """Row f-1 on device-resident data: torch uint8 [B,352] states -> packed records [B,stride]."""
        import torch

        st = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        _check(lib().hmpc_prepare_device(self._h, d_states.data_ptr(), B, dt_mpc, d_records.data_ptr(), ctypes.c_void_p(st)))

    def rollout_device(self, d_states, d_loop, B: int, ticks: int, d_wrench_log=None, d_record_log=None, stream=None,
                       dt_mpc: float = 0.04) -> None:
        """Row f-3: `ticks` closed-loop ticks (prepare -> solve -> advance) enqueued on one stream, no host in the
        loop.  torch CUDA tensors: states uint8 [B,352], loop uint8 [B,80], logs f32 [ticks,B,12] / uint8 [ticks,B,stride]."""
        import torch

        st = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        _check(lib().hmpc_rollout_device(self._h, d_states.data_ptr(), d_loop.data_ptr(), B, ticks, dt_mpc,
                                         d_wrench_log.data_ptr() if d_wrench_log is not None else None,
                                         d_record_log.data_ptr() if d_record_log is not None else None, ctypes.c_void_p(st)))

    # ---- multi-GPU: one process per GPU, batch sharded (hmpc_shard_*) ----
    @staticmethod
    def shard_unique_id() -> bytes:
        buf = ctypes.create_string_buffer(128)
        _check(lib().hmpc_shard_unique_id(buf))
        return buf.raw

    def shard_init(self, rank: int, world: int, unique_id: bytes) -> None:
        assert len(unique_id) == 128
        _check(lib().hmpc_shard_init(self._h, rank, world, ctypes.c_char_p(unique_id)))

    def solve_batch_sharded(self, records: np.ndarray, out, d_all=None, strict: bool = True):
        """This rank's slice through hmpc_solve_batch_sharded: results of the slice into `out` = (wrench f64 [b,12N], status
        i32 [b]) like solve_batch(out=...); `d_all` (torch CUDA f32 [world*b, 12N]) receives the one all-gather."""
        w, s = out
        rc = lib().hmpc_solve_batch_sharded(self._h, records.ctypes.data, len(records), w.ctypes.data, s.ctypes.data,
                                            ctypes.c_void_p(d_all.data_ptr()) if d_all is not None else None)
        _check(rc, allow_not_converged=not strict)
        return w, s

    def shard_wait(self) -> None:
        _check(lib().hmpc_shard_wait(self._h))

    def reset_warm_start(self, stream=None) -> None:
        """Forget the working sets the closed loop keeps between ticks (a new loop on this context starts cold)."""
        import torch

        st = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        _check(lib().hmpc_reset_warm_start(self._h, ctypes.c_void_p(st)))

    def swing_device(self, d_states, d_loop, d_phase, d_swing, B: int, d_cmd, dt: float = 0.001, dt_swing: float = 0.04,
                     stream=None) -> None:
        """Row f-4: one swingLegController::updateSwingLeg per robot on the device.  torch CUDA tensors: states uint8
        [B,352], loop uint8 [B,80], phase f64 [B], swing uint8 [B,72] (updated in place), cmd uint8 [B,232]."""
        import torch

        st = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        _check(lib().hmpc_swing_device(self._h, d_states.data_ptr(), d_loop.data_ptr(), d_phase.data_ptr(), d_swing.data_ptr(),
                                       B, dt, dt_swing, d_cmd.data_ptr(), ctypes.c_void_p(st)))

    def solve_device(self, d_records, B: int, d_wrench, d_status, stream=None) -> None:
        """Device-resident path.  Arguments are torch CUDA tensors (uint8 [B,stride], f32 [B,12N], i32 [B])."""
        import torch

        st = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        _check(lib().hmpc_solve_device(self._h, d_records.data_ptr(), B, d_wrench.data_ptr(), d_status.data_ptr(),
                                       ctypes.c_void_p(st)))

    def assemble_device(self, d_records, B: int, stream=None) -> dict:
        """Parity hook: un-reduced fp32 QP data of B packed records (torch tensors on the GPU)."""
        import torch

        N = self.horizon
        dev = torch.device("cuda", self.device)
        out = dict(
            H=torch.zeros((B, 12 * N, 12 * N), dtype=torch.float32, device=dev),
            g=torch.zeros((B, 12 * N), dtype=torch.float32, device=dev),
            Fblk=torch.zeros((B, 16, 12), dtype=torch.float32, device=dev),
            lb=torch.zeros((B, 16 * N), dtype=torch.float32, device=dev),
            ub=torch.zeros((B, 16 * N), dtype=torch.float32, device=dev),
        )
        st = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        _check(lib().hmpc_assemble_device(self._h, d_records.data_ptr(), B, out["H"].data_ptr(), out["g"].data_ptr(),
                                          out["Fblk"].data_ptr(), out["lb"].data_ptr(), out["ub"].data_ptr(),
                                          ctypes.c_void_p(st)))
        return out

    def close(self) -> None:
        if getattr(self, "_h", None):
            lib().hmpc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
