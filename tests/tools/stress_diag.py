import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from hector_simulation_b200 import interface, scenarios
from oracle import oracle_py as O
from oracle import qp_dual_active_set as G
from conftest import rel_err
N, B = 10, 8192
rng = np.random.default_rng(9001)
recs = np.zeros(B, dtype=scenarios.UPDATE_DTYPE)
for i in range(B):
    kind = i % 3
    table = scenarios.walking_table(N, int(rng.integers(0, N))) if kind == 0 else (scenarios.standing_table(N) if kind == 1 else (rng.random(2 * N) < 0.7).astype(np.int32))
    rpy = rng.normal(0.0, 0.15, 3); pos = np.array([0.0, 0.0, scenarios.BODY_HEIGHT]) + rng.normal(0.0, 0.06, 3); vx = rng.uniform(-1.0, 1.0)
    b = scenarios.boundary_inputs(pos, rpy, rng.normal(0, 0.3, 3) + [vx, 0, 0], rng.normal(0, 0.6, 3), rng.normal(0, 0.15, 10), table, N, v_des_body=(vx, 0.0), yaw_rate=rng.uniform(-0.5, 0.5), pos_des_err=rng.normal(0, 0.05, 2))
    scenarios.to_record(b, N, recs[i])
mpc = interface.BatchedMPC(B, N)
w, st = mpc.solve_batch(recs, strict=False)
print("codes", np.bincount(interface.status_code(st), minlength=5), "iters max", interface.status_iters(st).max())
idx = np.arange(0, B, 32)
setup = O.make_setup(N)
ref, info = O.solve_batch(recs[idx], setup)
e0 = rel_err(w[idx], ref, 12); ef = rel_err(w[idx], ref)
print("u0: max %.2e med %.2e | full: max %.2e med %.2e" % (e0.max(), np.median(e0), ef.max(), np.median(ef)), "nWSR max", info[:, 1].max())
# referee on the 6 worst
for k in np.argsort(-ef)[:6]:
    Q = O.reduced_qp(recs[idx[k]], setup)
    x, inf = G.solve(Q["H"], Q["g"], Q["A"], Q["lb"], Q["ub"], tol=1e-12, max_iter=3000)
    full = np.zeros(120); full[Q["var_ind"]] = x
    n = np.linalg.norm(full)
    print("inst", idx[k], "gpu-qp %.2e  gpu-referee %.2e  qp-referee %.2e  |x| %.1f iters gpu %d nWSR %d" % (ef[k], np.linalg.norm(w[idx[k]] - full) / n, np.linalg.norm(ref[k] - full) / n, n, interface.status_iters(st)[idx[k]], info[k, 1]))
