"""Static view of a kernel's SASS: instructions per source function block (by line ranges of the traversal headers), on the CPU box.
    python tools/sass_static.py voxelraymarcher_b200/libvrm_b200.so render_lean_kernelILi0ELi0 vrm_lean.cuh
Counts every instruction once under the OUTERMOST line of the named header in its inline chain."""
import collections
import re
import sys

sys.path.insert(0, __file__.rsplit("/", 1)[0])
from sass_hotspots import parse_sass, sass_with_lines  # noqa: E402


def main():
    so, kernel, header = sys.argv[1], sys.argv[2], sys.argv[3]
    parents = [int(v) for v in sys.argv[4:]]   # optional: descend below these lines of the header (outermost first)
    name, dis = sass_with_lines(so, kernel)
    sass = parse_sass(name, dis)
    by = collections.Counter()
    ops = collections.Counter()
    for off, text, chain in sass:
        key = "?"
        rev = list(reversed(chain))
        for i, (f, l) in enumerate(rev):
            if f == header:
                key = l
                j = i
                for want in parents:       # follow the requested parent lines inwards
                    if l == want and j + 1 < len(rev) and rev[j + 1][0] == header:
                        j += 1
                        l = rev[j][1]
                        key = l
                    else:
                        if l != want:
                            key = None
                        break
                break
        if key is None:
            continue
        by[key] += 1
        ops[re.sub(r"^@!?U?P\d\s+", "", text).split()[0].split(".")[0]] += 1
    print(name, len(sass), "instructions")
    for k in sorted(by, key=lambda v: (isinstance(v, str), v)):
        print(f"  {header}:{k}  {by[k]}")
    print("  opcodes:", ", ".join(f"{k} {v}" for k, v in ops.most_common(14)))


if __name__ == "__main__":
    main()
