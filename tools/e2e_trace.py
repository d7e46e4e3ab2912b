import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from hector_simulation_b200 import interface, scenarios
recs,_ = scenarios.make_batch(2, 1024)
mpc = interface.BatchedMPC(1024, 10)
for i in range(8): mpc.solve_batch(recs)
