"""Development check run on the GPU box: parity of assembly and of the solve, plus rough timings."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from hector_simulation_b200 import interface, scenarios  # noqa: E402
from oracle import oracle_py as O  # noqa: E402


def main():
    N = int(os.environ.get("HMPC_N", "10"))
    B = int(os.environ.get("HMPC_B", "512"))
    cfg = int(os.environ.get("HMPC_CFG", "3"))
    res = {}
    recs, _ = scenarios.make_batch(cfg, B, horizon=N)
    setup = O.make_setup(N)
    mpc = interface.BatchedMPC(max(B, 4096), N)
    res["classes"] = [mpc.class_config(0), mpc.class_config(1)]
    t = time.time()
    ref, info = O.solve_batch(recs, setup)
    res["oracle_ms_per_solve"] = (time.time() - t) / B * 1e3
    # ---- assembly parity (bitwise) ----
    packed = torch.from_numpy(interface.pack_records(recs, N)).cuda()
    nchk = min(B, 64)
    asm = mpc.assemble_device(packed, nchk)
    torch.cuda.synchronize()
    Hd = asm["H"].cpu().numpy(); gd = asm["g"].cpu().numpy(); Fd = asm["Fblk"].cpu().numpy()
    nbad_H = nbad_g = nbad_F = 0
    maxrel = 0.0
    for i in range(nchk):
        f = O.formulate_f32(recs[i], setup)
        iu = np.triu_indices(12 * N)
        a, b = Hd[i][iu], f["H"][iu]
        nbad_H += int((a.view(np.uint32) != b.view(np.uint32)).sum())
        maxrel = max(maxrel, float(np.abs(a - b).max() / np.abs(b).max()))
        nbad_g += int((gd[i].view(np.uint32) != f["g"].view(np.uint32)).sum())
        nbad_F += int((Fd[i].view(np.uint32) != f["Fblk"].view(np.uint32)).sum())
        assert np.array_equal(asm["lb"][i].cpu().numpy(), f["lb"]) and np.array_equal(asm["ub"][i].cpu().numpy(), f["ub"])
    res.update(asm_instances=nchk, H_bits_differ=nbad_H, g_bits_differ=nbad_g, F_bits_differ=nbad_F, H_maxrel=maxrel)
    # ---- solve parity ----
    wrench, status = mpc.solve_batch(recs, strict=False)
    code = interface.status_code(status)
    rel0 = np.linalg.norm(wrench[:, :12] - ref[:, :12], axis=1) / np.maximum(np.linalg.norm(ref[:, :12], axis=1), 1e-9)
    relf = np.linalg.norm(wrench - ref, axis=1) / np.maximum(np.linalg.norm(ref, axis=1), 1e-9)
    res.update(codes=np.bincount(code, minlength=5).tolist(), rel_u0_max=float(rel0.max()), rel_u0_med=float(np.median(rel0)),
               rel_full_max=float(relf.max()), iters_max=int(interface.status_iters(status).max()),
               iters_med=float(np.median(interface.status_iters(status))), nact_max=int(interface.status_nactive(status).max()),
               nwsr_max=int(info[:, 1].max()), nwsr_med=float(np.median(info[:, 1])),
               swing_zero=bool((wrench[ref == 0] == 0).all()))
    worst = int(np.argmax(rel0))
    res["worst"] = dict(i=worst, rel=float(rel0[worst]), iters=int(interface.status_iters(status)[worst]), nwsr=int(info[worst, 1]),
                        nv=int(info[worst, 2]))
    # ---- timing ----
    d_w = torch.zeros((B, 12 * N), dtype=torch.float32, device="cuda")
    d_s = torch.zeros(B, dtype=torch.int32, device="cuda")
    for _ in range(3):
        mpc.solve_device(packed, B, d_w, d_s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        mpc.solve_device(packed, B, d_w, d_s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    res.update(kernel_ms=ms, qp_per_s=B / ms * 1e3)
    # ---- per-stage latency (clock64 stamps of thread 0 of each CTA) ----
    import ctypes
    clk = torch.zeros((B, 32), dtype=torch.int64, device="cuda")
    interface.lib().hmpc_debug_set_clock_buffer(ctypes.c_void_p(clk.data_ptr()))
    mpc.solve_device(packed, B, d_w, d_s)
    torch.cuda.synchronize()
    interface.lib().hmpc_debug_set_clock_buffer(None)
    c = clk.cpu().numpy()
    dur = np.diff(c[:, :7], axis=1).astype(np.float64)
    names = ["load+prologue", "powers/M/d", "H,g chains", "sweep", "dual active set", "polish+scatter"]
    res["stage_cycles"] = {n: dict(med=float(np.median(dur[:, i])), p99=float(np.percentile(dur[:, i], 99))) for i, n in enumerate(names)}
    res["stage3_items_cycles_med"] = float(np.median(c[:, 7] - c[:, 2]))
    fine = {"s5 init (x0, slacks)": (4, 8), "s5 round0: slacks->smem": (8, 9), "s5 round0: candidates+slots": (9, 10),
            "s5 round0: T columns... S build": (10, 11), "s5 round0: S sweep": (11, 12), "s5 round0: multipliers+prune": (12, 13),
            "s5 round0: marks+barrier": (13, 14), "s5 round0: x, slacks": (14, 15), "s5 all block rounds": (8, 16), "s5 dual iteration after": (16, 5),
            "s4 step0 total": (3, 19), "s4 step1: W frags": (19, 20), "s4 step1: diag update+inverse": (20, 21), "s4 step1: rank-8 updates": (21, 22),
            "s4 step1: pivot row+transposed publish": (22, 23), "s4 step1: barrier wait": (23, 24)}
    res["fine_cycles_med"] = {k: float(np.median((c[:, b] - c[:, a])[(c[:, a] > 0) & (c[:, b] > 0)])) if ((c[:, a] > 0) & (c[:, b] > 0)).any() else None
                              for k, (a, b) in fine.items()}
    res["cta_total_cycles"] = dict(med=float(np.median(c[:, 6] - c[:, 0])), max=float((c[:, 6] - c[:, 0]).max()))
    it = interface.status_iters(d_s.cpu().numpy())
    gi = dur[:, 4]
    res["gi_cycles_per_iter_med"] = float(np.median(gi[it > 0] / it[it > 0]))
    t = time.time()
    for _ in range(5):
        mpc.solve_batch(recs, strict=False)
    res["e2e_ms"] = (time.time() - t) / 5 * 1e3
    res["e2e_qp_per_s"] = B / res["e2e_ms"] * 1e3
    print(json.dumps(res, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"gpu_check_cfg{cfg}_N{N}_B{B}.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
