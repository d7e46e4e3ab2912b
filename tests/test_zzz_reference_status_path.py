"""The reference boundary's status path (include/hector_mpc_b200.h, part 1): a tick that fails at run time keeps the previous
wrench and reports through hmpc_reference_last_rc() instead of ending the process.  In its own file, after all others: it
injects a failure into the library (in a subprocess) and exercises the abort policy."""
import os

import pytest

from conftest import ROOT


@pytest.mark.gpu
def test_reference_boundary_survives_a_failing_tick():
    """The reference's boundary has no error channel; a run-time failure under update_problem_data must not end a 1 kHz
    controller's process.  Fault injection: the library's debug hook makes the next host-buffer solve return HMPC_ERR_CUDA
    (what a CUDA failure looks like to the boundary).  The failing tick prints one line, get_solution keeps the previous
    wrench, hmpc_reference_last_rc() reports the error, the tick after it is normal again; HMPC_REFERENCE_ABORT=1 restores
    abort-on-failure.  (In a subprocess: the abort policy ends the process.)"""
    import subprocess
    import sys

    code = r'''
import sys
import numpy as np
sys.path.insert(0, %r)
from hector_simulation_b200 import interface, scenarios
b = scenarios.stand_inputs(10)
args = (b["p"], b["v"], b["q"], b["w"], b["r"], b["joint_angles"], b["yaw"], b["weights"], b["state_trajectory"], b["Alpha_K"], b["gait"])
interface.setup_problem(scenarios.DT_MPC, 10, scenarios.MU_PASSED, scenarios.F_MAX)
interface.update_problem_data(*args)
assert interface.reference_last_rc() == 0
s0 = [interface.get_solution(i) for i in range(12)]
assert abs(s0[2] - 47.84) < 0.05
interface.lib().hmpc_debug_fail_next_solves(1)
interface.update_problem_data(*args)          # fails inside the library
rc = interface.reference_last_rc()
s1 = [interface.get_solution(i) for i in range(12)]
print("RC", rc, "SAME", s1 == s0, flush=True)
interface.update_problem_data(*args)          # and the next tick is normal again
print("NEXT", interface.reference_last_rc(), [interface.get_solution(i) for i in range(12)] == s0, flush=True)
''' % ROOT
    env = dict(os.environ)
    env.pop("HMPC_REFERENCE_ABORT", None)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
    line = [l for l in r.stdout.splitlines() if l.startswith("RC ")][0].split()
    assert int(line[1]) == 2 and line[3] == "True", r.stdout          # HMPC_ERR_CUDA, previous wrench kept
    assert [l for l in r.stdout.splitlines() if l.startswith("NEXT ")][0].split()[1:] == ["0", "True"], r.stdout
    assert "keeping the previous solution" in r.stderr
    env["HMPC_REFERENCE_ABORT"] = "1"
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode != 0 and "RC " not in r.stdout      # aborted inside the failing tick
