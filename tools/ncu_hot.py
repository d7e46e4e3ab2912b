"""Hot source lines of a kernel in an .ncu-rep: samples per line split into barrier waits and everything else (the
critical path is in the "everything else" of the slowest warp), with the dominant stall reason and executed instructions.

    python tools/ncu_hot.py gpurun_out/prof.ncu-rep 'hmpc_solve_kernel<(int)128' ILi128ELi7ELi10ELi0 [top_n]
"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines as nl  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REASONS = ["stall_barrier", "stall_branch_resolving", "stall_dispatch", "stall_lg", "stall_long_sb", "stall_math", "stall_mio",
           "stall_no_inst", "stall_not_selected", "stall_selected", "stall_short_sb", "stall_wait", "stall_membar", "stall_misc"]


def main():
    rep, ksub, msub = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = nl.sass_rows(rep, ksub)
    lt = nl.line_table(os.path.join(ROOT, "hector_simulation_b200", "libhector_mpc_b200.so"), msub)
    src = open(os.path.join(ROOT, "hector_simulation_b200", "csrc", "hmpc_device.cuh")).read().splitlines()
    agg = collections.defaultdict(lambda: collections.Counter())
    for i, r in enumerate(rows[: len(lt)]):
        k = lt[i]
        if k is None:
            continue
        a = agg[k]
        a["inst"] += int(r["Instructions Executed"] or 0)
        for s in REASONS:
            a[s] += int(r.get(s) or 0)
        a["conf"] += int(r.get("L1 Wavefronts Shared Excessive") or 0)
    tot = collections.Counter()
    for a in agg.values():
        tot.update(a)
    nb = sum(tot[s] for s in REASONS) - tot["stall_barrier"]
    print(f"samples: {sum(tot[s] for s in REASONS)} (barrier {tot['stall_barrier']}), instructions {tot['inst']}, excessive smem wavefronts {tot['conf']}")
    print("totals by reason:", ", ".join(f"{s[6:]} {tot[s]}" for s in sorted(REASONS, key=lambda s: -tot[s]) if tot[s]))
    ranked = sorted(agg.items(), key=lambda kv: -(sum(kv[1][s] for s in REASONS) - kv[1]["stall_barrier"]))
    print(f"{'non-bar':>8} {'%':>5} {'barrier':>7} {'inst':>9} {'conf':>7}  top reason      line")
    for (f, l), a in ranked[:top]:
        work = sum(a[s] for s in REASONS) - a["stall_barrier"]
        rs = max((s for s in REASONS if s != "stall_barrier"), key=lambda s: a[s])
        text = src[l - 1].strip()[:100] if f == "hmpc_device.cuh" and 0 < l <= len(src) else f
        print(f"{work:8d} {100 * work / max(nb, 1):5.1f} {a['stall_barrier']:7d} {a['inst']:9d} {a['conf']:7d}  {rs[6:]:<14} {l:5d}: {text}")


if __name__ == "__main__":
    main()
