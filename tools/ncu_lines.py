"""Correlate an .ncu-rep's per-SASS-instruction samples with CUDA source lines.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep 'hmpc_solve_kernel<(int)128' [top_n] [mangled-substring e.g. ILi128E]

ncu's CLI source page carries no line column, so the kernel's cubin is disassembled with
`nvdisasm --print-line-info` and zipped with the SASS rows by instruction order.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_rows(rep, kernel_substr):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    res, hdr, take = [], None, False
    for r in rows:
        if len(r) >= 2 and r[0] == "Kernel Name":
            take = kernel_substr in r[1]
            hdr = None
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if take and hdr and len(r) == len(hdr):
            res.append(dict(zip(hdr, r)))
    return res


def line_table(lib, mangled_substr):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
    cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "--print-line-info", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
    lines = []
    cur_fn, cur_line, active = None, None, False
    for ln in dis.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            active = mangled_substr in m.group(1)
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            lines.append(cur_line)
    return lines


def main():
    rep, ksub = sys.argv[1], sys.argv[2]
    topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    msub = sys.argv[4] if len(sys.argv) > 4 else "hmpc_solve_kernel"  # substring of the mangled name, e.g. ILi128ELi7ELi3E
    rows = sass_rows(rep, ksub)
    lt = line_table(os.path.join(ROOT, "hector_simulation_b200", "libhector_mpc_b200.so"), msub)
    if len(rows) == 2 * len(lt):  # ncu prints the listing twice
        rows = rows[: len(lt)]
    print(f"sass rows {len(rows)}, disasm instrs {len(lt)}")
    agg = defaultdict(lambda: [0, 0])
    tot = 0
    for i, r in enumerate(rows):
        key = lt[i] if i < len(lt) else None
        s = int(r.get("# Samples", "0") or 0)
        e = int(r.get("Instructions Executed", "0") or 0)
        agg[key][0] += s
        agg[key][1] += e
        tot += s
    src = open(os.path.join(ROOT, "hector_simulation_b200", "csrc", "hmpc_device.cuh")).read().splitlines()
    print(f"total samples {tot}")
    for key, (s, e) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
        text = src[key[1] - 1].strip()[:100] if key and key[0] == "hmpc_device.cuh" and key[1] <= len(src) else ""
        print(f"{100.0 * s / max(tot, 1):6.2f}%  samples={s:7d} inst={e:9d}  {key}  {text}")


if __name__ == "__main__":
    main()
