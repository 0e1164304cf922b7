"""CPU tier: the PRODUCT's traversal core (voxelraymarcher_b200/csrc/vrm_core.cuh -- the source the sm_100a kernels are
built from) compiled for the host by tests/hostsim, checked against the oracle.  Catches arithmetic / control-flow
divergence without a GPU; the GPU tier repeats the comparison through the C ABI on the real kernels."""
import numpy as np
import pytest

from tests.common import COMBOS, MINI_CAMERAS, PROBE_CAMERAS, build_oracle, camera, lookup_queries, oracle_kind, po, scenes


@pytest.fixture(params=["nested", "flat"], autouse=True)
def traversal_form(request):
    """Both forms of the traversal the kernels are built from: vrm_core.cuh (nested loops) and vrm_flat.cuh (state machine)."""
    po._lib("sim").sim_set_flat(1 if request.param == "flat" else 0)
    yield request.param
    po._lib("sim").sim_set_flat(1)


@pytest.mark.parametrize("storage,algo", COMBOS)
def test_core_matches_oracle_probe(storage, algo):
    kind = oracle_kind()
    po.set_lighting(kind)
    po.set_lighting("sim")
    xyz, rgb = scenes.probe_scene()
    a, b = build_oracle(kind, xyz, rgb, storage), build_oracle("sim", xyz, rgb, storage)
    assert a.info() == b.info()
    q = lookup_queries(xyz, 5000, seed=5)
    for x, y in zip(a.lookup(q), b.lookup(q)):
        assert np.array_equal(x, y)
    for o, l, fov in PROBE_CAMERAS:
        cam = camera(o, l, fov, 256, 144, kind)
        ra = a.render(cam, 256, 144, algo, scale=8, want_counters=True, want_lookups=True)
        rb = b.render(cam, 256, 144, algo, scale=8, want_counters=True, want_lookups=True)
        for k in ("rgb", "hits", "counters", "lookups"):
            assert np.array_equal(ra[k], rb[k]), (storage, algo, k)


@pytest.mark.parametrize("storage,algo", COMBOS)
def test_core_matches_oracle_incoherent_rays(storage, algo):
    kind = oracle_kind()
    po.set_lighting(kind)
    po.set_lighting("sim")
    xyz, rgb = scenes.sparse_shells(256, 64, seed=7, fill_pct=50)
    a, b = build_oracle(kind, xyz, rgb, storage), build_oracle("sim", xyz, rgb, storage)
    rays = scenes.random_rays(30000, (130.0, 97.0, 121.0), seed=42)
    ta, tb = a.trace_rays(rays, algo, want_counters=True), b.trace_rays(rays, algo, want_counters=True)
    for k in ("colour", "hits", "counters"):
        assert np.array_equal(ta[k], tb[k]), (storage, algo, k)


@pytest.mark.parametrize("kw", [dict(use_point=True, position=(60.0, 90.0, 80.0)), dict(use_shadows=False),
                                dict(direction=(0.2, 0.9, -0.38), colour=(1.0, 0.8, 0.6))])
def test_core_lighting_variants(kw):
    kind = oracle_kind()
    xyz, rgb = scenes.mini_scene()
    try:
        if "direction" in kw:
            kw = dict(kw, direction=po.unit_vector(kw["direction"], kind))
        po.set_lighting(kind, **kw)
        po.set_lighting("sim", **kw)
        for storage, algo in COMBOS:
            a, b = build_oracle(kind, xyz, rgb, storage), build_oracle("sim", xyz, rgb, storage)
            for o, l, fov in MINI_CAMERAS:
                cam = camera(o, l, fov, 128, 72, kind)
                ra, rb = a.render(cam, 128, 72, algo), b.render(cam, 128, 72, algo)
                assert np.array_equal(ra["rgb"], rb["rgb"]) and np.array_equal(ra["hits"], rb["hits"]), (storage, algo, kw)
    finally:
        po.set_lighting(kind)
        po.set_lighting("sim")


def axis_aligned_rays():
    """Rays with exactly-zero direction components (+0 and -0), from inside and outside the scene: the unguarded divisions
    of the reference produce inf / NaN here (SURVEY.md §7 hard part 3 'NaN/Inf behaviour')."""
    dirs = []
    for axis in range(3):
        for sgn in (1.0, -1.0):
            for z in (0.0, -0.0):
                d = [z, z, z]
                d[axis] = sgn
                dirs.append(d)
    dirs += [[0.6, 0.8, 0.0], [0.6, -0.8, -0.0], [0.0, 0.6, -0.8], [-0.0, -0.6, 0.8], [0.8, 0.0, 0.6], [-0.8, -0.0, -0.6]]
    origins = [(40.3, 33.7, 36.2), (-30.5, 12.25, -70.75), (100.5, 40.5, 200.5), (32.0, 32.0, 32.0), (20.5, 300.5, -40.5)]
    rays = [list(o) + d for o in origins for d in dirs]
    return np.array(rays, np.float32)


@pytest.mark.parametrize("storage,algo", COMBOS)
def test_core_axis_aligned_rays_terminate_and_match(storage, algo):
    kind = oracle_kind()
    po.set_lighting(kind)
    po.set_lighting("sim")
    xyz, rgb = scenes.probe_scene()
    a, b = build_oracle(kind, xyz, rgb, storage), build_oracle("sim", xyz, rgb, storage)
    rays = axis_aligned_rays()
    ta, tb = a.trace_rays(rays, algo, threads=1), b.trace_rays(rays, algo, threads=1)
    assert np.array_equal(ta["colour"], tb["colour"])
    assert np.array_equal(ta["hits"], tb["hits"])


def crawl_scene_and_rays(n=1500, seed=5):
    """Rays that get stuck on a cluster face of an empty VCS cluster and advance by EPSILON per skip iteration in the
    reference (a small negative direction component, |d| < ulp(position) / (2 EPSILON)): three floors on the coordinate
    planes, rays skimming above them.  The oracle executes ~5e7 cluster-skip iterations for these 3000 rays; the product
    fast-forwards them (crawl_skip, vrm_core.cuh) and must stay bit-identical, event counters included."""
    xyz, rgb = scenes.checker_floor(0, 64, 0, 64, y=0)
    xyz = np.concatenate([xyz, xyz[:, [1, 2, 0]], xyz[:, [2, 0, 1]]])
    rgb = np.concatenate([rgb, rgb, rgb])
    rng = np.random.default_rng(seed)
    o = np.stack([rng.uniform(33, 63.9, n), rng.uniform(20, 60, n), rng.uniform(0.5, 8, n)], 1)
    d = np.stack([-rng.uniform(0.004, 0.018, n), -rng.uniform(0.1, 0.4, n), rng.uniform(0.8, 1, n)], 1)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([o, d], 1)
    rolled = rays.copy()
    rolled[:, 0:3] = np.roll(rays[:, 0:3], 1, axis=1)
    rolled[:, 3:6] = np.roll(rays[:, 3:6], 1, axis=1)
    return xyz.astype(np.int32), rgb, np.concatenate([rays, rolled]).astype(np.float32)


@pytest.mark.parametrize("algo", ["original", "longestaxis"])
def test_core_crawl_fast_forward_is_bit_exact(algo):
    kind = oracle_kind()
    po.set_lighting(kind)
    po.set_lighting("sim")
    xyz, rgb, rays = crawl_scene_and_rays()
    a, b = build_oracle(kind, xyz, rgb, "vcs"), build_oracle("sim", xyz, rgb, "vcs")
    ta, tb = a.trace_rays(rays, algo, want_counters=True), b.trace_rays(rays, algo, want_counters=True)
    assert int(ta["counters"][0]) > 10_000_000          # the reference really does crawl here
    for k in ("colour", "hits", "counters"):
        assert np.array_equal(ta[k], tb[k]), k


def test_exact_reciprocal_divisions_of_the_primary_ray():
    """primary_ray_flat replaces (x + 0.5) / W, (H - y + 0.5) / H and rel / len by reciprocal + FMA-residual divisions: they must
    be the IEEE quotients, bit for bit, for every pixel coordinate of the image sizes in use and for random normalisations."""
    import ctypes as C
    lib = po._lib("sim")
    lib.sim_check_image_division.restype = C.c_uint64
    lib.sim_check_image_division.argtypes = [C.c_uint32]
    lib.sim_check_common_division.restype = C.c_uint64
    lib.sim_check_common_division.argtypes = [C.c_uint32, C.c_uint64]
    for n in (1, 2, 3, 7, 72, 128, 144, 180, 256, 320, 360, 540, 640, 720, 960, 1000, 1080, 1280, 1920, 2160, 3840, 4096, 7680, 65535, 1 << 20):
        assert lib.sim_check_image_division(n) == 0, n
    assert lib.sim_check_common_division(1, 20_000_000) == 0
