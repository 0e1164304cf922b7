"""Runs of consecutive SASS instructions with equal execution counts in a dump written by tools/sass_lines.py --dump: index, length,
executions per warp, threads per instruction, share of the kernel's warp instructions, innermost source line.

    python tools/sass_segments.py dump.txt <warps in the launch> [min share %]
"""
import sys


def main():
    path, warps = sys.argv[1], float(sys.argv[2])
    floor = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
    recs = []
    for line in open(path):
        p = line.rstrip("\n").split(None, 4)
        recs.append((p[0], int(p[1]), float(p[2]), int(p[3]), p[4] if len(p) > 4 else ""))
    tot = sum(r[1] for r in recs)
    print(f"total {tot}  per warp {tot / warps:.1f}")
    i = 0
    while i < len(recs):
        j = i
        while j + 1 < len(recs) and recs[j + 1][1] == recs[i][1]:
            j += 1
        n, w = j - i + 1, recs[i][1]
        share = 100.0 * n * w / tot
        if share >= floor:
            thr = sum(r[2] for r in recs[i:j + 1]) / n
            smp = sum(r[3] for r in recs[i:j + 1])
            src = recs[i][4].split("  ")[-1].strip()[:80]
            print(f"idx {i:5d} len {n:4d} exec/warp {w / warps:7.2f} thr {thr:5.1f} share {share:5.2f}% smp {smp:6d}  {src}")
        i = j + 1


if __name__ == "__main__":
    main()
