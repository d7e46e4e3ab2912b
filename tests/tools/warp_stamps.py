"""Per-warp clock stamps of one block step of stage 4: who reaches the step's barrier last, and what it did before.
Developer tool, run on the GPU box with a library built with the stamps compiled in:
    nvcc <__graft_entry__.NVCC_FLAGS> -DHMPC_WARP_STAMPS=<step> -o hector_simulation_b200/libhector_mpc_b200.so \
         hector_simulation_b200/csrc/hmpc_capi.cu
Result of round 2 (configs[1], steps 1 and 4): profiles/stage_cycles_r2.json "stage4_per_warp"."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from hector_simulation_b200 import interface, scenarios  # noqa: E402

B, N = 1024, 10
recs, _ = scenarios.make_batch(2, B, horizon=N)
mpc = interface.BatchedMPC(4096, N)
packed = torch.from_numpy(interface.pack_records(recs, N)).cuda()
d_w = torch.zeros((B, 12 * N), dtype=torch.float32, device="cuda")
d_s = torch.zeros(B, dtype=torch.int32, device="cuda")
for _ in range(3):
    mpc.solve_device(packed, B, d_w, d_s)
clk = torch.zeros((B, 32), dtype=torch.int64, device="cuda")
interface.lib().hmpc_debug_set_clock_buffer(ctypes.c_void_p(clk.data_ptr()))
mpc.solve_device(packed, B, d_w, d_s)
torch.cuda.synchronize()
interface.lib().hmpc_debug_set_clock_buffer(None)
c = clk.cpu().numpy().reshape(B, 4, 8)[:, :, :6].astype(np.float64)
t0 = c[:, :, 0].min(axis=1, keepdims=True)            # first warp to enter the step
names = ["enter", "W frags", "look-ahead+inverse", "rank-8 updates", "pivot row+publish", "after barrier"]
rel = c - t0[:, :, None]
print("median cycles since the first warp entered the step, per warp (rows) and event (columns):", names)
print(np.median(rel, axis=0).round())
seg = np.diff(c, axis=2)
print("median segment durations per warp:", names[1:])
print(np.median(seg, axis=0).round())
last = np.argmax(c[:, :, 4], axis=1)
print("warp arriving last at the barrier: histogram", np.bincount(last, minlength=4))
