"""CPU dry run of the closed loop (row f-3) with the ORACLE as the solver: checks that the numpy mirror of
hmpc_advance_kernel keeps walking robots upright before GPU time is spent on it.  Test infrastructure only.

    python tools/rollout_cpu_sim.py [robots] [ticks]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from hector_simulation_b200 import scenarios  # noqa: E402
from oracle import oracle_py as O  # noqa: E402
from test_state_prepare import _host_prepared  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    N = 10
    _, inputs = scenarios.make_batch(5, n, horizon=N)
    states, loop = scenarios.make_rollout(inputs, N, standing=None)
    setup = O.make_setup(N)
    for t in range(T):
        recs = _host_prepared(states, N)
        q, info = O.solve_batch(recs, setup)
        status = np.where(info[:, 0] == 0, 0, 1).astype(np.int32) | (info[:, 1].astype(np.int32) << 8)
        scenarios.advance_numpy(states, loop, q, status, N)
        if t % 20 == 19 or t == T - 1:
            print("t=%3d  z %.3f..%.3f  |rp| max %.3f  x %.2f..%.2f  vx %.2f..%.2f  fail %d  nWSR mean %.1f" % (
                t + 1, states["position"][:, 2].min(), states["position"][:, 2].max(), np.abs(states["rpy"][:, :2]).max(),
                states["position"][:, 0].min(), states["position"][:, 0].max(), states["vWorld"][:, 0].min(),
                states["vWorld"][:, 0].max(), loop["failures"].sum(), loop["iters_total"].sum() / loop["ticks"].sum()))


if __name__ == "__main__":
    main()
