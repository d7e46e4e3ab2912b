"""Regenerates the golden fixtures in this directory (run where /root/reference exists).

    python tests/golden/make_golden.py

Each cfg*_h*.npz holds synthetic robot states of one BASELINE.json config (generator and seeds:
hector_simulation_b200/scenarios.py) together with what the ORACLE returned for them:

  records   uint8 [n, 3016]   the reference's `update_data_t` records, byte for byte
  q_soln    f64   [n, 12N]    restated solve_mpc (fp32 formulation) + the reference's own qpOASES 3.2,
                              called as SolverMPC.cpp:702-712 does  -> what get_solution(i) would return
  info      i32   [n, 4]      {return code, nWSR, reduced variables, reduced constraints}
  H,g,Fblk,lb,ub  (first 4 records) the un-reduced fp32 QP data of the restated formulation

The reference itself ships no tests or vectors, so these are outputs of the reference's
solver on the restated formulation in its CANONICAL arithmetic (double trig narrowed to float — what the CUDA kernel
reproduces bit for bit).  The restatement is pinned against the reference's own sources compiled unchanged
(make_ref_compiled.py -> ref_compiled_h10.npz holds that build's outputs for the same records; oracle/solve_mpc_oracle.cpp
header explains the two trig modes).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from hector_simulation_b200 import scenarios  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

CASES = [  # (name, cfg, batch, horizon)
    ("cfg1_h10", 1, 1, 10),
    ("cfg2_h10", 2, 64, 10),
    ("cfg3_h10", 3, 64, 10),
    ("cfg4_h5", 4, 16, 5),
    ("cfg4_h16", 4, 16, 16),
]


def main():
    assert O.has_qpoases(), "needs the reference's qpOASES (build oracle/ where /root/reference exists)"
    for name, cfg, batch, N in CASES:
        recs, _ = scenarios.make_batch(cfg, batch, horizon=N)
        setup = O.make_setup(N)
        q, info = O.solve_batch(recs, setup)
        assert (info[:, 0] == 0).all(), (name, info[:, 0])
        nf = min(4, batch)
        forms = [O.formulate_f32(recs[i], setup) for i in range(nf)]
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            records=recs.view(np.uint8).reshape(batch, -1), q_soln=q, info=info, horizon=np.int32(N), cfg=np.int32(cfg),
            H=np.stack([f["H"] for f in forms]), g=np.stack([f["g"] for f in forms]),
            Fblk=np.stack([f["Fblk"] for f in forms]), lb=np.stack([f["lb"] for f in forms]), ub=np.stack([f["ub"] for f in forms]),
        )
        print(name, "nWSR", info[:, 1].min(), info[:, 1].max(), "nv", np.unique(info[:, 2]))


# degenerate_zero_force_h10.npz is not generated here: its record was captured from the closed-loop harness
# (tools/closed_loop_debug.py, a falling robot) and its solution computed with the same oracle calls as above.


if __name__ == "__main__":
    main()
