"""Static SASS instruction count of one solve-kernel instantiation by source stage (nvdisasm -g line info).

    python tools/sass_stage_hist.py [kernel-name-substring]      (default: the horizon-10 class-0 kernel)
"""
import bisect
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "hector_simulation_b200", "libhector_mpc_b200.so")
SRC = os.path.join(ROOT, "hector_simulation_b200", "csrc", "hmpc_device.cuh")


def main():
    sub = sys.argv[1] if len(sys.argv) > 1 else "hmpc_solve_kernelILi128ELi7ELi10ELi0"
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=td, check=True, stdout=subprocess.DEVNULL)
        cubin = os.path.join(td, [f for f in os.listdir(td) if f.endswith(".cubin")][0])
        dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True, check=True).stdout
    cnt, ops = collections.Counter(), collections.Counter()
    infun, cur = False, None
    for ln in dis.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            infun = sub in m.group(1)
            continue
        if not infun:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = int(m.group(2)) if m.group(1).endswith("hmpc_device.cuh") else -1
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", ln)
        if m:
            cnt[cur] += 1
            ops[m.group(1).split(".")[0]] += 1
    src = open(SRC).read().splitlines()
    marks = [(i + 1, l.strip()) for i, l in enumerate(src)
             if "// ----------------" in l or l.startswith("__device__") or l.startswith("__global__") or l.startswith("template")]
    starts = [m[0] for m in marks]
    agg = collections.Counter()
    for line, c in cnt.items():
        if line is None or line < 0:
            agg["other"] += c
            continue
        j = bisect.bisect_right(starts, line) - 1
        agg[f"{marks[j][0]:5d} {marks[j][1][:100]}"] += c
    tot = sum(agg.values())
    print(f"{sub}: {tot} SASS instructions")
    for k, v in sorted(agg.items()):
        print(f"{v:6d} {100 * v / max(tot, 1):5.1f}%  {k}")
    print("opcodes:", ", ".join(f"{k} {v}" for k, v in ops.most_common(24)))


if __name__ == "__main__":
    main()
