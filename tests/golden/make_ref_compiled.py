"""Regenerates tests/golden/ref_compiled_h10.npz (run where /root/reference exists).

    python tests/golden/make_ref_compiled.py

Outputs of THE REFERENCE'S OWN formulation sources — hector_control/ConvexMPC/{SolverMPC,RobotState,
convexMPC_interface}.cpp compiled unchanged by oracle/Makefile against oracle/eigen_shim (a stand-in for the absent
Eigen) and the reference's qpOASES, called through `resize_qp_mats` + `solve_mpc` — on the records of the
cfg1/cfg2/cfg3 horizon-10 fixtures in this directory:

  <cfg>_q       f64 [n, 120]     what get_solution(i) returns after the reference's solve
  <cfg>_H/g/A/lb/ub/x0  (first 2 records of cfg2, cfg3) the reference's file-scope qH, qg, fmat, L_b, U_b, x_0

The reference's trig calls resolve to libm's FLOAT functions (its TU pulls <math.h> in through qpOASES' Utils.ipp:36),
whose last bits depend on the glibc build and CPU; tests therefore hold other machines to these vectors with a few-ulp /
1e-5 tolerance and require bit identity only against a libref_mpc.so built on the same machine.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import load_golden  # noqa: E402
from oracle import oracle_py as O  # noqa: E402


def main():
    assert O.has_reference_build(), "needs oracle/_ref/libref_mpc.so (make -C oracle where /root/reference exists)"
    setup = O.make_setup(10)
    out = {}
    for name in ("cfg1", "cfg2", "cfg3"):
        g = load_golden(name + "_h10")
        q, F = O.ref_solve(g["records"], setup, formulation=True)
        out[name + "_q"] = q
        if name != "cfg1":
            for k in ("H", "g", "A", "lb", "ub", "x0"):
                out[name + "_" + k] = F[k][:2]
        rel = np.linalg.norm(q[:, :12] - g["q_soln"][:, :12], axis=1) / np.linalg.norm(g["q_soln"][:, :12], axis=1)
        print(name, "first-step gap of the canonical-arithmetic fixture to the compiled reference: max %.2e median %.2e" % (rel.max(), np.median(rel)))
    np.savez_compressed(os.path.join(HERE, "ref_compiled_h10.npz"), **out)


if __name__ == "__main__":
    main()
