import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from hector_simulation_b200 import interface, scenarios
from oracle import oracle_py as O
from test_closed_loop import _step
from conftest import rel_err
N, B, T = 10, 192, 30
rng = np.random.default_rng(505)
setup = O.make_setup(N)
mpc = interface.BatchedMPC(B, N)
states, feet, joints, phase = [], [], [], []
for i in range(B):
    b = scenarios._random_state(rng, N, scenarios.walking_table(N, i % N), moving=True)
    rpy = scenarios.quat_to_rpy(b["q"])
    states.append((rpy, b["p"].copy(), b["w"].copy() * 0.2, b["v"].copy() * 0.2))
    feet.append(b["p_foot"].copy()); joints.append(rng.normal(0, 0.05, 10)); phase.append(i % N)
recs = np.zeros(B, dtype=scenarios.UPDATE_DTYPE)
worst = (0, None)
for t in range(T):
    for i in range(B):
        rpy, p, w, v = states[i]
        table = scenarios.standing_table(N) if i % 4 == 0 else scenarios.walking_table(N, (phase[i] + t) % N)
        b = scenarios.boundary_inputs(p, rpy, v, w, joints[i], table, N, feet_world=feet[i])
        scenarios.to_record(b, N, recs[i])
    wrench, status = mpc.solve_batch(recs, strict=False)
    ref, info = O.solve_batch(recs, setup)
    e = rel_err(wrench, ref, 12)
    k = int(np.argmax(e))
    print("tick", t, "max rel", "%.2e" % e[k], "inst", k, "iters", interface.status_iters(status)[k], "nWSR", info[k, 1], "code", interface.status_code(status)[k], "|u0|", "%.2f" % np.linalg.norm(ref[k, :12]), "z", "%.3f" % states[k][1][2], "n>1e-4:", int((e > 1e-4).sum()))
    if e[k] > worst[0] and t <= 14:
        worst = (e[k], recs[k].copy(), wrench[k].copy(), ref[k].copy())
    for i in range(B):
        states[i] = _step(states[i], wrench[i, :12], feet[i], scenarios.DT_MPC)
np.save(os.path.join(ROOT, "gpurun_out", "worst_record.npy"), np.frombuffer(worst[1].tobytes(), dtype=np.uint8))
np.save(os.path.join(ROOT, "gpurun_out", "worst_gpu.npy"), worst[2]); np.save(os.path.join(ROOT, "gpurun_out", "worst_ref.npy"), worst[3])
