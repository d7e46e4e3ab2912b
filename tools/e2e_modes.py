"""Times the host-buffer path (hmpc_solve_batch) under the copy modes / chunk counts selected by environment
variables, one subprocess per mode (the library reads them once).    python tools/e2e_modes.py [batch]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, time, json, os
sys.path.insert(0, %r)
import numpy as np
from hector_simulation_b200 import interface, scenarios
B = int(sys.argv[1])
recs, inputs = scenarios.make_batch(2, B)
mpc = interface.BatchedMPC(B, 10)
w = np.zeros((B, 120)); s = np.zeros(B, np.int32)
if os.environ.get("HMPC_PIN") == "1": mpc.pin(recs, w, s)
for _ in range(20): mpc.solve_batch(recs, out=(w, s))
lat = []
for _ in range(300):
    t = time.perf_counter(); mpc.solve_batch(recs, out=(w, s)); lat.append(time.perf_counter() - t)
lat = np.array(lat) * 1e3
states = scenarios.make_states(inputs, 10)
for _ in range(20): mpc.solve_batch_states(states, out=(w, s))
lat2 = []
for _ in range(300):
    t = time.perf_counter(); mpc.solve_batch_states(states, out=(w, s)); lat2.append(time.perf_counter() - t)
lat2 = np.array(lat2) * 1e3
b = scenarios.stand_inputs(10)
tick = []
for _ in range(300):
    t = time.perf_counter()
    interface.setup_problem(0.04, 10, 0.25, 500.0)
    interface.update_problem_data(b["p"], b["v"], b["q"], b["w"], b["r"], b["joint_angles"], b["yaw"], b["weights"], b["state_trajectory"], b["Alpha_K"], b["gait"])
    u0 = [interface.get_solution(i) for i in range(12)]
    tick.append(time.perf_counter() - t)
tick = np.array(tick[50:]) * 1e3
w2, s2 = mpc.solve_batch(recs)
print(json.dumps({"ms_p50": float(np.percentile(lat, 50)), "ms_mean": float(lat.mean()), "qps": B / lat.mean() * 1e3,
                  "states_ms_mean": float(lat2.mean()), "tick_ms_p50": float(np.percentile(tick, 50)),
                  "checksum": float(np.abs(w2).sum()), "bad": int((interface.status_code(s2) != 0).sum())}))
''' % ROOT


def main():
    B = sys.argv[1] if len(sys.argv) > 1 else "1024"
    for mode in ({"HMPC_ZEROCOPY": "0", "HMPC_CHUNKS": "2"}, {"HMPC_ZEROCOPY": "1", "HMPC_CHUNKS": "1"}, {}, {"HMPC_PIN": "1"}):
        env = dict(os.environ, **mode)
        r = subprocess.run([sys.executable, "-c", CHILD, B], env=env, capture_output=True, text=True, timeout=600)
        print("%s  %s %s" % (mode or "default", r.stdout.strip(), r.stderr.strip()[-300:]), flush=True)


if __name__ == "__main__":
    main()
