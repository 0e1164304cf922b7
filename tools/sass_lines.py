"""Per-source-line issue-slot accounting from a source-page CSV saved on the GPU box (tools/gpu_capture.sh: <tag>_sass.csv.gz) and the
object file (or .so) of the SAME build: joins the capture's per-instruction counters with `nvdisasm -gi` line tables.

    python tools/sass_lines.py gpurun_out/r02f_ncu_render_vcs_longestaxis_sass.csv.gz build/vrm_render.o render_kernelILi0ELi0ELb0ELi2 [--top 40] [--dump out.txt]
"""
import argparse
import collections
import csv
import gzip
import os
import re
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sass_hotspots import parse_sass  # noqa: E402


def disasm(obj, kernel_substr):
    tmp = tempfile.mkdtemp(prefix="sass_")
    if obj.endswith(".cubin"):
        cubins = [os.path.abspath(obj)]
    else:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
        cubins = [os.path.join(tmp, f) for f in sorted(os.listdir(tmp))]
    for cubin in cubins:
        elf = subprocess.run(["cuobjdump", "-elf", cubin], stdout=subprocess.PIPE, text=True).stdout
        for line in elf.splitlines():
            m = re.match(r"\s*(0x[0-9a-f]+)\s+\S+\s+0x[0-9a-f]+\s+0x2\s+0x10\s+\S+\s+(\S+)", line)
            if m and kernel_substr in m.group(2) and not m.group(2).startswith("."):
                idx, name = m.group(1), m.group(2)
                return name, subprocess.run(["nvdisasm", "-gi", "-fun", idx, cubin], stdout=subprocess.PIPE, text=True).stdout
    raise SystemExit(f"kernel containing {kernel_substr!r} not found in {obj}")


def counts(path):
    rows = list(csv.reader(gzip.open(path, "rt")))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    col = {n: i for i, n in enumerate(rows[h])}
    out = []
    for r in rows[h + 1:]:
        if len(r) < len(rows[h]) or not r[0].startswith("0x"):
            continue
        out.append(dict(sass=r[col["Source"]].strip(), warp=int(r[col["Instructions Executed"]]), thread=int(r[col["Thread Instructions Executed"]]), samples=int(r[col["# Samples"]])))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv"); ap.add_argument("obj"); ap.add_argument("kernel")
    ap.add_argument("--top", type=int, default=40); ap.add_argument("--dump")
    a = ap.parse_args()
    name, dis = disasm(a.obj, a.kernel)
    sass = parse_sass(name, dis)
    cnt = counts(a.csv)
    n = min(len(sass), len(cnt))
    op = lambda t: re.sub(r"^@!?U?P\d\s+", "", t).split()[0].split(".")[0]
    bad = sum(1 for i in range(n) if op(sass[i][1]) != op(cnt[i]["sass"]))
    print(f"{name}: {len(sass)} SASS instructions in the object, {len(cnt)} in the capture, {bad} opcode mismatches")
    tw = sum(c["warp"] for c in cnt[:n]); tt = sum(c["thread"] for c in cnt[:n])
    print(f"{tw / 1e6:.1f} M warp-instr, {tt / max(tw, 1):.2f} threads/instr")
    by_line, thr_line = collections.Counter(), collections.Counter()
    for i in range(n):
        _, text, chain = sass[i]
        c = cnt[i]
        # the innermost line that lies in the traversal headers / the kernel file
        key = next((f"{f}:{l}" for f, l in chain if f in ("vrm_flat.cuh", "vrm_core.cuh", "vrm_render.cu", "vrm_lean.cuh")), "?")
        by_line[key] += c["warp"]; thr_line[key] += c["thread"]
    print("\n-- by innermost product line: warp-instr %, threads/instr")
    for k, v in by_line.most_common(a.top):
        print(f"  {k:26s} {100 * v / tw:6.2f} %  {thr_line[k] / max(v, 1):5.1f}")
    if a.dump:
        with open(a.dump, "w") as f:
            for i in range(n):
                off, text, chain = sass[i]
                c = cnt[i]
                f.write(f"{off:05x} {c['warp']:>10d} {c['thread'] / max(c['warp'], 1):5.1f} {c['samples']:>5d}  {text:64s} {' < '.join(f'{x}:{y}' for x, y in chain)}\n")


if __name__ == "__main__":
    main()
