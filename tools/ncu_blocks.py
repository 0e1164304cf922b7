"""Read an .ncu-rep (one kernel, --set full --import-source on) and print: totals, stall mix, and the SASS basic blocks that
execute the most warp instructions (with their average active threads).  Writes the annotated SASS to <out>.

    python tools/ncu_blocks.py gpurun_out/prof.ncu-rep /tmp/annot.txt [top]
"""
import csv
import subprocess
import sys


def main():
    rep, out = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    if rep.endswith(".csv.gz"):     # the source page as saved on the GPU box by tools/gpu_capture.sh
        import gzip
        txt = gzip.open(rep, "rt").read()
    else:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    print(rows[0])
    hdr, data = rows[start], rows[start + 1:]
    ix = {h: i for i, h in enumerate(hdr)}
    g = lambda r, k: int(r[ix[k]]) if r[ix[k]] not in ("", "-") else 0
    tot = sum(g(r, "Instructions Executed") for r in data)
    thr = sum(g(r, "Thread Instructions Executed") for r in data)
    smp = sum(g(r, "# Samples") for r in data)
    print(f"static {len(data)}  warp-instructions {tot}  threads/instr {thr / tot:.2f}  samples {smp}")
    stalls = {h: sum(g(r, h) for r in data) for h in hdr if h.startswith("stall_") and "Not Issued" not in h}
    print("stalls:", ", ".join(f"{k[6:]} {100 * v / smp:.1f}%" for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:9]))
    blocks, cur = [], None
    with open(out, "w") as f:
        for i, r in enumerate(data):
            n, t, s = g(r, "Instructions Executed"), g(r, "Thread Instructions Executed"), g(r, "# Samples")
            f.write(f"{i:5d} {n:10d} {t / n if n else 0:5.1f} {s:5d}  {r[ix['Source']].strip()}\n")
            if cur and cur["n"] == n:
                cur["len"] += 1; cur["s"] += s; cur["t"] += t
            else:
                cur = dict(start=i, n=n, len=1, s=s, t=t); blocks.append(cur)
    for b in sorted(blocks, key=lambda b: -b["n"] * b["len"])[:top]:
        if b["n"] == 0:
            break
        print(f"start {b['start']:5d} len {b['len']:4d} exec {b['n']:10d} share {100 * b['n'] * b['len'] / tot:5.2f}% thr {b['t'] / (b['n'] * b['len']):5.1f} samples {100 * b['s'] / smp:5.2f}%")


if __name__ == "__main__":
    main()
