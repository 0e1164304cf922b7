"""Where do a kernel's issue slots go?  Joins the per-instruction counters of an ncu capture (`--set full --import-source on`)
with the line table of the cubin (`nvdisasm -gi`) and prints warp-instructions / thread-instructions per source line, per
"frame" (the outermost vrm_flat.cuh / vrm_core.cuh line of the inline chain) and per opcode.  Runs on the CPU box.

    python tools/sass_hotspots.py gpurun_out/prof.ncu-rep voxelraymarcher_b200/libvrm_b200.so render_kernelILi0ELi0ELb0ELb1 [--top 40]

The .so must be the build that was profiled (the SASS offsets are matched one to one; the tool checks the opcodes agree)."""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_with_lines(so_path, kernel_substr):
    tmp = tempfile.mkdtemp(prefix="sass_")
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so_path)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    for cubin in sorted(os.listdir(tmp)):
        elf = subprocess.run(["cuobjdump", "-elf", os.path.join(tmp, cubin)], stdout=subprocess.PIPE, text=True).stdout
        for line in elf.splitlines():
            m = re.match(r"\s*(0x[0-9a-f]+)\s+\S+\s+0x[0-9a-f]+\s+0x2\s+0x10\s+\S+\s+(\S+)", line)
            if m and kernel_substr in m.group(2) and not m.group(2).startswith("."):
                idx, name = m.group(1), m.group(2)
                dis = subprocess.run(["nvdisasm", "-gi", "-fun", idx, os.path.join(tmp, cubin)], stdout=subprocess.PIPE, text=True).stdout
                return name, dis
    raise SystemExit(f"kernel containing {kernel_substr!r} not found in {so_path}")


def parse_sass(name, dis):
    """-> list of (offset, opcode text, [(file, line), ...] innermost first)"""
    out, chain, pending = [], [], []
    in_text = False
    for line in dis.splitlines():
        if line.startswith("\t.section\t.text.") or line.startswith(".text."):
            in_text = name in line
        if not in_text:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
        if m:
            pending.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            if pending:
                # nvdisasm -gi prints the chain innermost first, then repeats the callers; dedupe keeping order
                seen, chain = set(), []
                for p in pending:
                    if p not in seen:
                        seen.add(p); chain.append(p)
                pending = []
            out.append((int(m.group(1), 16), m.group(2).strip(), list(chain)))
    return out


def ncu_counts(rep, kernel_substr):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {n: i for i, n in enumerate(hdr)}
    out = []
    base = None
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr) or not r[0].startswith("0x"):
            continue
        a = int(r[0], 16)
        base = a if base is None else base
        out.append(dict(off=a - base, sass=r[col["Source"]].strip(), warp=int(r[col["Instructions Executed"]]), thread=int(r[col["Thread Instructions Executed"]]),
                        samples=int(r[col["# Samples"]]), sectors=int(r[col["L2 Theoretical Sectors Global"]] or 0)))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep"); ap.add_argument("so"); ap.add_argument("kernel")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--dump", help="write the annotated listing here")
    a = ap.parse_args()
    name, dis = sass_with_lines(a.so, a.kernel)
    sass = parse_sass(name, dis)
    cnt = ncu_counts(a.rep, a.kernel)
    if len(sass) != len(cnt):
        print(f"warning: {len(sass)} SASS instructions in the .so vs {len(cnt)} in the capture", file=sys.stderr)
    n = min(len(sass), len(cnt))
    bad = sum(1 for i in range(n) if sass[i][1].split()[0].lstrip("@!P0123456789T ") [:4] != cnt[i]["sass"].split()[0].lstrip("@!P0123456789T ")[:4])
    if bad > n // 50:
        print(f"warning: {bad} of {n} opcodes differ -- is this the profiled build?", file=sys.stderr)
    total_w = sum(c["warp"] for c in cnt[:n]); total_t = sum(c["thread"] for c in cnt[:n]); total_s = sum(c["samples"] for c in cnt[:n])
    print(f"{name}\n{n} instructions, {total_w / 1e6:.1f} M warp-instr, {total_t / 1e6:.1f} M thread-instr, {total_t / max(total_w, 1):.2f} threads/instr, {total_s} samples")

    def frame(chain):
        # outermost line that lies in the traversal headers (what block of the state machine the instruction belongs to)
        for f, l in reversed(chain):
            if f in ("vrm_flat.cuh", "vrm_core.cuh"):
                return f"{f}:{l}"
        return f"{chain[-1][0]}:{chain[-1][1]}" if chain else "?"

    by_line, by_frame, by_op = collections.Counter(), collections.Counter(), collections.Counter()
    thr_line, thr_frame, smp_frame = collections.Counter(), collections.Counter(), collections.Counter()
    for i in range(n):
        off, text, chain = sass[i]
        c = cnt[i]
        inner = f"{chain[0][0]}:{chain[0][1]}" if chain else "?"
        by_line[inner] += c["warp"]; thr_line[inner] += c["thread"]
        fr = frame(chain)
        by_frame[fr] += c["warp"]; thr_frame[fr] += c["thread"]; smp_frame[fr] += c["samples"]
        op = re.sub(r"^@!?U?P\d\s+", "", text).split()[0].split(".")[0]
        by_op[op] += c["warp"]
    print("\n-- by frame (outermost traversal-header line): warp-instr %, threads/instr, stall samples %")
    for k, v in by_frame.most_common(a.top):
        print(f"  {k:26s} {100 * v / total_w:6.2f} %  {thr_frame[k] / max(v, 1):5.1f}  {100 * smp_frame[k] / max(total_s, 1):6.2f} %")
    print("\n-- by innermost line")
    for k, v in by_line.most_common(a.top):
        print(f"  {k:26s} {100 * v / total_w:6.2f} %  {thr_line[k] / max(v, 1):5.1f}")
    print("\n-- by opcode")
    for k, v in by_op.most_common(25):
        print(f"  {k:12s} {100 * v / total_w:6.2f} %")
    if a.dump:
        with open(a.dump, "w") as f:
            for i in range(n):
                off, text, chain = sass[i]
                c = cnt[i]
                f.write(f"{off:05x} {c['warp']:>10d} {c['thread'] / max(c['warp'], 1):5.1f} {c['samples']:>5d}  {text:60s} {' < '.join(f'{a_}:{b_}' for a_, b_ in chain)}\n")


if __name__ == "__main__":
    main()
