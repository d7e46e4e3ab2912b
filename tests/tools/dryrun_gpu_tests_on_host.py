"""Developer tool: run the bodies of -m gpu tests on the CPU with the device calls redirected to the host emulation of the
kernel source (tests/host_emul/, built by `pytest tests/test_kernel_source_on_host.py`), to catch plumbing mistakes — field
names, shapes, tolerances — before a test meets a GPU.  Not part of any test run; the product is untouched (the fake
BatchedMPC below lives in this script only).

    python -m pytest tests/test_kernel_source_on_host.py -q -k prepare     # builds tests/host_emul/_build/
    python tests/tools/dryrun_gpu_tests_on_host.py
"""
import sys, ctypes, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from hector_simulation_b200 import interface
import test_kernel_source_on_host as KS

L = ctypes.CDLL(os.path.join(KS.BUILD, "libkernel_source_on_host.so"))
_p = KS._p

class FakeMPC:
    def __init__(self, B, N, device=0, **kw): self.B, self.horizon = B, N
    def prepare_device(self, d_states, B, d_rec, stream=None, dt_mpc=0.04):
        st = np.ascontiguousarray(d_states.numpy()); out = d_rec.numpy()
        tmp = np.ascontiguousarray(out.copy())
        L.emul_prepare(_p(st), B, self.horizon, ctypes.c_double(dt_mpc), _p(tmp)); out[:] = tmp
    def swing_device(self, d_states, d_loop, d_phase, d_swing, B, d_cmd, dt=0.001, dt_swing=0.04, stream=None):
        st = np.ascontiguousarray(d_states.numpy()); lo = np.ascontiguousarray(d_loop.numpy()); ph = np.ascontiguousarray(d_phase.numpy())
        sw = d_swing.numpy(); cmd = d_cmd.numpy()
        assert sw.flags.c_contiguous and cmd.flags.c_contiguous
        L.emul_swing(_p(st), _p(lo), _p(ph), _p(sw), B, self.horizon, ctypes.c_double(dt), ctypes.c_double(dt_swing), _p(cmd))
    def solve_batch_torques(self, records, strict=True):
        w, st, tau, _, _ = KS._solve(L, records, self.horizon)
        return w, tau, st
    def solve_batch(self, records, strict=True, out=None):
        w, st, _, _, _ = KS._solve(L, records, self.horizon, tau=False)
        return w, st
    def close(self): pass

interface.BatchedMPC = FakeMPC
torch.Tensor.cuda = lambda self, *a, **k: self
torch.cuda.synchronize = lambda *a, **k: None
_full, _zeros, _tensor = torch.full, torch.zeros, torch.tensor
def _cpu(f):
    def g(*a, **k):
        k.pop('device', None); return f(*a, **k)
    return g
torch.full, torch.zeros, torch.tensor = _cpu(_full), _cpu(_zeros), _cpu(_tensor)

import test_zz_device_vs_reference_vectors as Z
Z.test_device_preparation_reproduces_the_reference_controllers_records(); print("zz prepare ok")
Z.test_device_swing_controller_follows_the_reference_controller(); print("zz swing ok")
Z.test_device_solve_and_torques_follow_the_reference_controller(); print("zz solve+torques ok")
# drop-in test body with the reference's own solver standing in for the GPU library
from oracle import oracle_py as O
_RC = O.ReferenceController
O.ReferenceController = lambda dt, it, drop_in=False: _RC(dt, it, drop_in=False)
O.has_reference_tick_dropin = lambda: True
Z.test_reference_controller_runs_on_the_gpu_library(); print("zz drop-in loop ok (reference solver standing in)")
import test_gpu_parity as G
for name in ("cfg1", "cfg2", "cfg3"):
    G.test_wrench_vs_compiled_reference_vectors(None, name)
print("gpu_parity compiled-reference vectors ok")
