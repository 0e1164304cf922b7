"""Generates the committed golden fixtures from the UNMODIFIED reference built for the host
(oracle/_ref/libvrm_ref_host.so, see oracle/ref_host.cpp).  Needs /root/reference, so it only runs in the build
container:   python tests/golden/make_golden.py
The reference ships no golden vectors of its own; these pin the oracle and the CUDA path to the
reference's behaviour on the stand-in scenes."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests.common import COMBOS, GOLDEN_DIR, MINI_CAMERAS, PROBE_CAMERAS, build_oracle, camera, lookup_queries, po, scenes  # noqa: E402

W, H = 160, 90


def main():
    assert po.available("refh"), "build oracle/_ref first (make -C oracle ref)"
    po.set_lighting("refh")
    out = {}
    for name, (xyz, rgb), scale in (("probe", scenes.probe_scene(), 8), ("mini", scenes.mini_scene(), 1)):
        q = lookup_queries(xyz, 4000, seed=7)
        out[f"{name}_queries"] = q
        for storage in ("hashtable", "vcs"):
            ref = build_oracle("refh", xyz, rgb, storage)
            info = ref.info()
            out[f"{name}_{storage}_info"] = np.array([info["diameter"], info["min_coord"], info["filled"]], np.int64)
            val, ex = ref.lookup(q)
            out[f"{name}_{storage}_lookup"] = val
            out[f"{name}_{storage}_exists"] = ex
            cams = PROBE_CAMERAS if name == "probe" else MINI_CAMERAS
            for ci, (o, l, fov) in enumerate(cams):
                cam = camera(o, l, fov, W, H, "refh")
                out[f"{name}_cam{ci}"] = cam
                for algo in ("original", "longestaxis"):
                    r = ref.render(cam, W, H, algo, scale=scale)
                    out[f"{name}_{storage}_{algo}_cam{ci}_rgb"] = r["rgb"]
                    out[f"{name}_{storage}_{algo}_cam{ci}_hits"] = r["hits"]
            ref.close()
    # lighting variants on the probe scene (point light on / shadows off), reference camera
    xyz, rgb = scenes.probe_scene()
    cam = camera(*PROBE_CAMERAS[1], W, H, "refh")
    for tag, kw in (("point", dict(use_point=True, position=(60.0, 90.0, 80.0))), ("noshadow", dict(use_shadows=False))):
        po.set_lighting("refh", **kw)
        for storage, algo in COMBOS:
            ref = build_oracle("refh", xyz, rgb, storage)
            r = ref.render(cam, W, H, algo, scale=8)
            out[f"probe_{storage}_{algo}_{tag}_rgb"] = r["rgb"]
            ref.close()
    po.set_lighting("refh")
    np.savez_compressed(os.path.join(GOLDEN_DIR, "reference_golden.npz"), **out)
    print("wrote", os.path.join(GOLDEN_DIR, "reference_golden.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
