"""Per-stage split of an ncu capture of hmpc_solve_kernel (samples by source-line range)."""
import bisect
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines as nl  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, ksub, msub = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = nl.sass_rows(rep, ksub)
    lt = nl.line_table(os.path.join(ROOT, "hector_simulation_b200", "libhector_mpc_b200.so"), msub)
    if len(rows) == 2 * len(lt):
        rows = rows[: len(lt)]
    src = open(os.path.join(ROOT, "hector_simulation_b200", "csrc", "hmpc_device.cuh")).read().splitlines()
    marks = [(i + 1, l.strip()) for i, l in enumerate(src)
             if "// ----------------" in l or l.startswith("__device__") or l.startswith("__global__")]
    starts = [m[0] for m in marks]
    agg, tot, toti = {}, 0, 0
    for i, r in enumerate(rows):
        k = lt[i] if i < len(lt) else None
        s = int(r["# Samples"] or 0)
        e = int(r["Instructions Executed"] or 0)
        tot += s
        toti += e
        if k is None or k[0] != "hmpc_device.cuh":
            key = "other"
        else:
            j = bisect.bisect_right(starts, k[1]) - 1
            key = f"{starts[j]:4d} {marks[j][1][:80]}" if j >= 0 else "pre"
        a = agg.setdefault(key, [0, 0])
        a[0] += s
        a[1] += e
    print(f"sass rows {len(rows)} / disasm {len(lt)}; samples {tot}; warp instructions {toti}")
    for k, (s, e) in sorted(agg.items()):
        print(f"{100 * s / max(tot, 1):6.2f}% samples  {100 * e / max(toti, 1):6.2f}% inst ({e:10d})  {k}")


if __name__ == "__main__":
    main()
