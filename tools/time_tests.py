import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from hector_simulation_b200 import interface
from conftest import load_golden
for name in ("cfg3_h10", "cfg4_h16", "degenerate_zero_force_h10"):
    g = load_golden(name); N = g["horizon"]
    t = time.time(); mpc = interface.BatchedMPC(len(g["records"]), N); t1 = time.time()
    w, st = mpc.solve_batch(g["records"], strict=False); t2 = time.time()
    w, st = mpc.solve_batch(g["records"], strict=False); t3 = time.time()
    print(name, "create %.2fs first solve %.3fs second %.4fs" % (t1 - t, t2 - t1, t3 - t2), "codes", np.bincount(interface.status_code(st)), "iters max", interface.status_iters(st).max(),
          [mpc.class_config(i) for i in range(3) if i < 3 and (i < 2 or True)][-1] if True else "")
    mpc.close()
