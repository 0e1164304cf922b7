"""A/B on a GPU box: time A/B builds of libvrm_b200.so (voxelraymarcher_b200/variants/libvrm_<name>.so, built with
`make -C voxelraymarcher_b200/csrc variant NAME=<name> VARIANT_FLAGS=...`) on the bench workload through tools/explore.py.

    python tools/ab_variants.py main old:mode=2 cta128 mb3@vcs:longestaxis

A spec is  <variant>[:mode=<VRM_RENDER_MODE>][:shadow=<VRM_SHADOW_FORM>][@<storage>:<algo>,...] ; `main` is the product library.  Each spec runs in its own
process (the library path is read at import).  Prints one line per (spec, combination) and writes gpurun_out/ab.json."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    specs = sys.argv[1:] or ["main"]
    rows = []
    for spec in specs:
        combos = None
        if "@" in spec:
            spec_l, combos = spec.split("@", 1)
        else:
            spec_l = spec
        parts = spec_l.split(":")
        name = parts[0]
        env = dict(os.environ)
        for kv in parts[1:]:
            k, v = kv.split("=")
            if k == "mode":
                env["VRM_RENDER_MODE"] = v
            elif k == "shadow":
                env["VRM_SHADOW_FORM"] = v
        if name != "main":
            env["VRM_B200_LIB"] = os.path.join(ROOT, "voxelraymarcher_b200", "variants", f"libvrm_{name}.so")
        out = os.path.join(ROOT, "gpurun_out", f"ab_{spec.replace(':', '_').replace('@', '_').replace(',', '_').replace('=', '')}.json")
        cmd = [sys.executable, os.path.join(ROOT, "tools", "explore.py"), "--iters", "9", "--out", out]
        if combos:
            cmd += ["--combos", combos]
        proc = subprocess.run(cmd, env=env, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if proc.returncode != 0:
            print(f"{spec}: FAILED\n{proc.stdout[-2000:]}", flush=True)
            continue
        for r in json.load(open(out)):
            rows.append(dict(spec=spec, storage=r["storage"], algo=r["algo"], ms=r["ms"], mrays=r["mrays"], hit_fraction=r["hit_fraction"]))
            print(f"{spec:28s} {r['storage']:9s} {r['algo']:11s} {r['ms']:8.3f} ms {r['mrays']:9.0f} Mrays/s  hit {r['hit_fraction']:.6f}", flush=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "ab.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
