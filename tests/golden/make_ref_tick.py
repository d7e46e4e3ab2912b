"""Regenerates tests/golden/ref_tick_walk.npz (run where /root/reference exists).

    python tests/golden/make_ref_tick.py

Per-tick outputs (`reftick_out_t`, oracle/ref_tick_probe.cpp) of THE REFERENCE'S OWN walking controller —
ConvexMPCLocomotion, GaitGenerator, LegController, SwingLegController, FootSwingTrajectory, DesiredCommand and the MPC
formulation files compiled unchanged by oracle/Makefile against oracle/eigen_shim, plus the reference's qpOASES — ticked
through the pose sequence of tests/test_reference_tick.py (520 ticks of the walking gait, 104 MPC solves).
The inputs are not stored: the test regenerates them from the same closed-form sequence.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import test_reference_tick as T  # noqa: E402
from oracle import oracle_py as O  # noqa: E402


def main():
    assert O.has_reference_tick(), "needs oracle/_ref/libref_tick.so (make -C oracle where /root/reference exists)"
    ticks = np.array(list(T.reference_ticks(O)), dtype=O.REFTICK_DTYPE)
    np.savez_compressed(os.path.join(HERE, "ref_tick_walk.npz"), ticks=ticks.view(np.uint8).reshape(len(ticks), -1))
    print(len(ticks), "ticks,", int(ticks["mpc_ran"].sum()), "MPC solves")
    # ref_tick_cases.npz: the shorter cases of test_reference_tick.CASES (zero command under the walking gait, standing gait)
    more = {}
    for case in T.CASES:
        if case == "walk":
            continue
        t = np.array(list(T.reference_ticks(O, case)), dtype=O.REFTICK_DTYPE)
        more[case] = t.view(np.uint8).reshape(len(t), -1)
        print(case, len(t), "ticks,", int(t["mpc_ran"].sum()), "MPC solves")
    np.savez_compressed(os.path.join(HERE, "ref_tick_cases.npz"), **more)


if __name__ == "__main__":
    main()
