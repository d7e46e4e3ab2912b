"""Regenerates tests/golden/stress_referee.npz (needs the oracle with qpOASES: run where /root/reference exists).

    python tests/golden/make_stress_fixture.py

Robustness workload (hector_simulation_b200.scenarios.make_stress_batch): states far outside the operating envelope under
walking / standing / random contact tables, 60-110 active rows at the optimum.  Stored per set: the records, qpOASES'
answer through the oracle (the reference's own solver, its own termination tolerance) and the answer of the tight-tolerance
fp64 referee (oracle/qp_dual_active_set.py, tol 1e-12) on the same canonical QP.  On these ill-conditioned problems qpOASES
itself is off the exact optimum by up to 6e-3 of the first-step wrench; the tests hold the CUDA path to the REFEREE and
check that, where the two CPU answers differ, the kernel agrees with the exact one."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from hector_simulation_b200 import scenarios  # noqa: E402
from oracle import oracle_py as O  # noqa: E402
from oracle import qp_dual_active_set as G  # noqa: E402

SETS = {"h10_x4": (10, 32, 4.0, 2), "h10_x8": (10, 32, 8.0, 3), "h14_x4": (14, 16, 4.0, 5)}   # horizon, batch, scale, seed
# one record beyond the conditioning limit of the kernel's sweep inversion (a robot lying on its side: max_i H_ii (H^-1)_ii =
# 2.9e5, cond(H) = 3e8): record 34 of make_stress_batch(40, 10, 8.0, 15).  qpOASES (Cholesky-based) still solves it.
LYING = ("h10_lying", 10, 40, 8.0, 15, [34])


def main():
    assert O.has_qpoases()
    out = {}
    todo = [(name, N, scenarios.make_stress_batch(B, N, scale, seed)) for name, (N, B, scale, seed) in SETS.items()]
    todo.append((LYING[0], LYING[1], scenarios.make_stress_batch(LYING[2], LYING[1], LYING[3], LYING[4])[LYING[5]]))
    for name, N, recs in todo:
        B = len(recs)
        setup = O.make_setup(N)
        q, info = O.solve_batch(recs, setup)
        ref = np.zeros((B, 12 * N))
        for k in range(B):
            Q = O.reduced_qp(recs[k], setup)
            x, inf = G.solve(Q["H"], Q["g"], Q["A"], Q["lb"], Q["ub"], tol=1e-12, max_iter=5000)
            assert inf["status"] == 0, (name, k)
            ref[k, Q["var_ind"]] = x
        d = np.linalg.norm(q[:, :12] - ref[:, :12], axis=1) / np.maximum(np.linalg.norm(ref[:, :12], axis=1), 1e-9)
        print("%s: %d records, qpOASES ok %d, qpOASES vs referee first step: median %.1e, worst %.1e" %
              (name, B, int((info[:, 0] == 0).sum()), np.median(d), d.max()))
        out[name + "_records"] = recs.view(np.uint8).reshape(B, -1)
        out[name + "_qpoases"] = q
        out[name + "_qpoases_ok"] = (info[:, 0] == 0)
        out[name + "_referee"] = ref
    np.savez_compressed(os.path.join(HERE, "stress_referee.npz"), **out)


if __name__ == "__main__":
    main()
