"""CPU tier: the PRODUCT's traversal core (voxelraymarcher_b200/csrc/vrm_core.cuh -- the source the sm_100a kernels are
built from) compiled for the host by tests/hostsim, checked against the oracle.  Catches arithmetic / control-flow
divergence without a GPU; the GPU tier repeats the comparison through the C ABI on the real kernels."""
import numpy as np
import pytest

from tests.common import COMBOS, MINI_CAMERAS, PROBE_CAMERAS, build_oracle, camera, lookup_queries, oracle_kind, po, scenes


FORMS = {"nested": 0, "flat": 1, "lean": 2, "fast": 3}


@pytest.fixture(params=list(FORMS), autouse=True)
def traversal_form(request):
    """The three forms of the traversal the kernels are built from: vrm_core.cuh (nested loops), vrm_flat.cuh (generic state
    machine) and vrm_lean.cuh (the hot kernels' state machine; rays it parks are re-traced by the generic one, as on the GPU)."""
    po._lib("sim").sim_set_flat(FORMS[request.param])
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


def pingpong_scene_and_rays():
    """The reference's second crawl pathology (vrm_flat.cuh pingpong_skip): longest-axis cluster jumps of a ray that sits exactly on
    a REGION face with one short axis and exactly on a cluster face with the other, both with direction components too small to
    move it -- it changes region twice per cycle and advances EPSILON per iteration.  Two x two regions with a wall far down
    the path, rays starting on the region face y = 64 / cluster face z = 24 (and the axis-rolled twins)."""
    ys, zs = np.meshgrid(np.arange(0, 128, dtype=np.int32), np.arange(0, 128, dtype=np.int32), indexing="ij")
    wall = np.stack([np.full(ys.size, 61, np.int32), ys.ravel(), zs.ravel()], 1)
    seeds = np.array([[1, 1, 1], [1, 70, 1], [1, 1, 70], [1, 70, 70]], np.int32)      # make all four regions exist near the start
    xyz = np.concatenate([wall, seeds])
    rgb = (np.arange(xyz.shape[0], dtype=np.uint32) * np.uint32(2654435761)) & np.uint32(0xFFFFFF) | np.uint32(0x010101)
    base = [((2.5, 64.0, 24.0), (0.99982786, -0.017536791, -0.0060571153)),
            ((3.25, 64.0, 48.0), (0.999665618, -0.0187616404, -0.0177927297)),
            ((5.125, 64.0, 56.0), (0.999815047, -0.0185558796, -0.00505277468)),
            ((2.5, 24.0, 64.0), (0.99982786, -0.0060571153, -0.017536791))]
    rays = [list(o) + list(d) for o, d in base]
    full_xyz, full_rgb, full_rays = [xyz], [rgb], [np.array(rays, np.float32)]
    for shift in (1, 2):                       # the same geometry with the axes rolled: every axis takes every role
        full_xyz.append(np.roll(xyz, shift, axis=1) + np.array([0, 0, 0], np.int32) + 256 * shift)
        full_rgb.append(rgb)
        r = np.array(rays, np.float32)
        r[:, 0:3] = np.roll(r[:, 0:3], shift, axis=1) + 256 * shift
        r[:, 3:6] = np.roll(r[:, 3:6], shift, axis=1)
        full_rays.append(r)
    return np.concatenate(full_xyz).astype(np.int32), np.concatenate(full_rgb), np.concatenate(full_rays).astype(np.float32)


def test_core_region_face_pingpong_fast_forward_is_bit_exact():
    import ctypes as C
    kind = oracle_kind()
    po.set_lighting(kind)
    po.set_lighting("sim")
    xyz, rgb, rays = pingpong_scene_and_rays()
    a, b = build_oracle(kind, xyz, rgb, "vcs"), build_oracle("sim", xyz, rgb, "vcs")
    lib = po._lib("sim")
    lib.sim_crawl_skipped.restype = C.c_ulonglong
    lib.sim_crawl_skipped()
    ta, tb = a.trace_rays(rays, "longestaxis", want_counters=True), b.trace_rays(rays, "longestaxis", want_counters=True)
    skipped = int(lib.sim_crawl_skipped())
    assert int(ta["counters"][0]) > 5_000_000           # the reference really does crawl here (twice per cycle across a region face)
    for k in ("colour", "hits", "counters"):
        assert np.array_equal(ta[k], tb[k]), k
    assert (ta["hits"][:, 3] == 1).all()                # every ray ends on the wall
    if po._lib("sim").sim_is_flat():
        assert skipped > 0.9 * int(ta["counters"][0])   # ... and the state machine fast-forwarded nearly all of it


def test_core_crawl_fast_forward_handles_exact_ties():
    """EPSILON * d / ulp(position) exactly k + 0.5: the additions round to the even mantissa, a constant even step once the
    mantissa is even (found on the 2048^3 orbit: 190 000 iterations for one pixel).  Same numbers as that pixel."""
    import ctypes as C
    kind = oracle_kind()
    po.set_lighting(kind)
    po.set_lighting("sim")
    ys, zs = np.meshgrid(np.arange(0, 64, dtype=np.int32), np.arange(0, 64, dtype=np.int32), indexing="ij")
    xyz = np.stack([np.full(ys.size, 62, np.int32), ys.ravel(), zs.ravel()], 1)
    rgb = np.full(xyz.shape[0], 0x80C0F0, np.uint32)
    d = (0.848770142, -0.0162127428, -0.528513312)
    # a cluster skip towards the y = 48 face lands exactly on it (the EPSILON overshoot rounds away); x in [16, 32): EPSILON * d.x = 44.5 ulp
    rays = np.array([[17.0, 48.02, 46.0, *d], [17.000002, 48.02, 46.0, *d], [9.0, 48.02, 46.0, *d], [33.0, 48.02, 46.0, *d]], np.float32)
    a, b = build_oracle(kind, xyz, rgb, "vcs"), build_oracle("sim", xyz, rgb, "vcs")
    lib = po._lib("sim")
    lib.sim_crawl_skipped.restype = C.c_ulonglong
    lib.sim_crawl_skipped()
    for algo in ("original", "longestaxis"):
        ta, tb = a.trace_rays(rays, algo, want_counters=True), b.trace_rays(rays, algo, want_counters=True)
        for k in ("colour", "hits", "counters"):
            assert np.array_equal(ta[k], tb[k]), (algo, k)
        if algo == "original":
            assert int(ta["counters"][0]) > 300_000     # ~100 000 iterations per ray in the reference
            assert int(lib.sim_crawl_skipped()) > 0.9 * int(ta["counters"][0])


@pytest.mark.parametrize("algo", ["original", "longestaxis"])
def test_core_rays_starting_on_power_of_two_coordinates(algo):
    """Regression: crawl_skip took an axis sitting on a cluster face at a power-of-two coordinate (local 16.0) for stuck although a
    negative EPSILON * d between a quarter and half an ulp DOES move it (the floats below a power of two are twice as dense), and
    fast-forwarded 80 000 iterations the reference never executes.  Rays from (80, 150, 80) = region-local (16, 22, 16)."""
    kind = oracle_kind()
    po.set_lighting(kind)
    po.set_lighting("sim")
    xyz, rgb = scenes.terrain(160, 77)
    a, b = build_oracle(kind, xyz, rgb, "vcs"), build_oracle("sim", xyz, rgb, "vcs")
    rays = scenes.random_rays(20000, (80.0, 150.0, 80.0), seed=101)
    ta, tb = a.trace_rays(rays, algo, want_counters=True), b.trace_rays(rays, algo, want_counters=True)
    for k in ("colour", "hits", "counters"):
        assert np.array_equal(ta[k], tb[k]), k


@pytest.mark.parametrize("storage,algo", COMBOS)
def test_core_scene_translation(storage, algo):
    """VoxelSceneInfo's translation vector (Main.cu:215 always passes zero, the routines take any): ray origins are moved by
    -translation before scaling and the lighting positions by +translation (Renderer.cuh:338-341, 821-822)."""
    kind = oracle_kind()
    po.set_lighting(kind)
    po.set_lighting("sim")
    xyz, rgb = scenes.probe_scene()
    a, b = build_oracle(kind, xyz, rgb, storage), build_oracle("sim", xyz, rgb, storage)
    o, l, fov = PROBE_CAMERAS[1]
    cam = camera(o, l, fov, 160, 90, kind)
    for tr in ((3.5, -2.25, 7.0), (-0.3, 0.7, 0.1)):
        ra = a.render(cam, 160, 90, algo, scale=8, translation=tr, want_counters=True, want_lookups=True)
        rb = b.render(cam, 160, 90, algo, scale=8, translation=tr, want_counters=True, want_lookups=True)
        assert (ra["hits"][..., 3] != 0).sum() > 1000
        for k in ("rgb", "hits", "counters", "lookups"):
            assert np.array_equal(ra[k], rb[k]), (storage, algo, tr, k)


@pytest.mark.parametrize("storage,algo", COMBOS)
def test_core_origins_a_few_ulps_around_cluster_faces(storage, algo):
    """Rays from origins whose coordinates sit a few ulps below / above a cluster face or an integer (tools/stress_modes.py `ulps` in
    small): the first steps round differently on either side of such a value.  Colour, hit voxel and event counters."""
    po.set_lighting("orc")
    po.set_lighting("sim")
    xyz, rgb = scenes.terrain(160, 77)
    a, b = build_oracle("orc", xyz, rgb, storage), build_oracle("sim", xyz, rgb, storage)
    f32 = np.float32
    up, down = (lambda v: float(np.nextafter(f32(v), f32(np.inf)))), (lambda v: float(np.nextafter(f32(v), f32(-np.inf))))
    origins = [(down(72.0), 150.0, up(up(40.0))), (up(16.0), down(down(down(136.0))), 81.0), (down(33.0), up(140.0), down(8.0))]
    for i, org in enumerate(origins):
        rays = scenes.random_rays(4000, org, seed=40 + i)
        ta, tb = a.trace_rays(rays, algo, want_counters=True), b.trace_rays(rays, algo, want_counters=True)
        for k in ("colour", "hits", "counters"):
            assert np.array_equal(ta[k], tb[k]), (storage, algo, org, k)


def coordinate64_rays():
    """Rays from (128, 64, 0) * 1/8 of the probe scene (scale 8): a corner shared by eight regions.  Their first EPSILON step leaves the
    region at -tiny on the two short axes and is rebased to exactly 64.0f, so the longest-axis walk tests voxels with a coordinate of 64
    before its grid values are back inside the region (found by tools/stress_diff.py, seed 77)."""
    bits = [(0xbbd13b05, 0xbba7a053, 0xbf7ffdce), (0xbc37ff18, 0xbc19b7c9, 0xbf7ff8fb), (0xbc76f148, 0xbb6a7051, 0xbf7ff823),
            (0xbf7ff978, 0xbbf7f442, 0xbc435e2e), (0xbc6cd52e, 0xbacde9b9, 0xbf7ff912), (0xbf7ff3e6, 0xbc26e79f, 0xbc8582e4)]
    rays = np.zeros((len(bits), 6), np.float32)
    rays[:, :3] = (16.0, 8.0, -0.0)
    rays[:, 3:] = np.array(bits, np.uint32).view(np.float32)
    return rays


@pytest.mark.parametrize("storage", ["vcs", "hashtable"])
def test_core_lookups_with_a_coordinate_of_64_are_empty(storage):
    """A lookup with a region-local coordinate of exactly 64 matches no stored voxel in the reference (its key x << 20 | y << 10 | z
    differs from every stored key), whatever cluster the overflowing bits alias into.  Six-bit packed codes would alias into voxel
    (.., 0, ..) of the neighbouring cluster row and these rays would hit voxel (127, 64, -1), which the reference passes."""
    po.set_lighting("orc")
    po.set_lighting("sim")
    xyz, rgb = scenes.probe_scene()
    a, b = build_oracle("orc", xyz, rgb, storage), build_oracle("sim", xyz, rgb, storage)
    rays = coordinate64_rays()
    ta, tb = a.trace_rays(rays, "longestaxis", scale=8, want_counters=True), b.trace_rays(rays, "longestaxis", scale=8, want_counters=True)
    assert not ta["hits"][:, 3].any()
    # (event counters only for the hash table: the VCS exists test of the reference is undefined when the coordinate of 64 is x)
    for k in ("colour", "hits") + (("counters",) if storage == "hashtable" else ()):
        assert np.array_equal(ta[k], tb[k]), k


@pytest.mark.parametrize("storage,algo", [("hashtable", "longestaxis"), ("hashtable", "original"), ("vcs", "original")])
def test_core_render_from_region_corner_cameras(storage, algo):
    """Cameras exactly on a corner shared by eight regions, looking along the faces (tests/test_parity_gpu.py has the same case on the
    kernels): everything the hash table and the original algorithm do there is defined in the reference, so rgb, hit voxels and every
    event counter must match."""
    kind = oracle_kind()
    po.set_lighting(kind)
    po.set_lighting("sim")
    xyz, rgb = scenes.probe_scene()
    a, b = build_oracle(kind, xyz, rgb, storage), build_oracle("sim", xyz, rgb, storage)
    for o, l in (((16.0, 8.0, 0.0), (15.9, 7.9, -10.0)), ((8.0, 16.0, 8.0), (7.95, 0.0, 7.9))):
        cam = camera(o, l, 60.0, 192, 108, kind)
        ra = a.render(cam, 192, 108, algo, scale=8, want_counters=True, want_lookups=True)
        rb = b.render(cam, 192, 108, algo, scale=8, want_counters=True, want_lookups=True)
        for k in ("rgb", "hits", "counters", "lookups"):
            assert np.array_equal(ra[k], rb[k]), (storage, algo, o, k)


def test_crawl_skip_equals_the_iterations_it_replaces():
    """Brute force: 60 000 pseudo-random crawl situations (positions on cluster faces / integers / powers of two and a few ulps around
    them; EPSILON steps from "cannot move" to hundreds of ulps).  Whenever crawl_skip fast-forwards, executing the skipped iterations
    one by one must give the same bits and stay inside the skipped cluster cell.  (Two defects of the first version -- a coordinate
    on a power of two taken for stuck, and a step landing exactly on a binade's first float -- were found by differential runs
    against the oracle and are covered here.)"""
    import ctypes as C
    lib = po._lib("sim")
    lib.sim_check_crawl_skip.restype = C.c_uint64
    lib.sim_check_crawl_skip.argtypes = [C.c_uint32, C.c_uint64, C.POINTER(C.c_uint64)]
    used = C.c_uint64()
    assert lib.sim_check_crawl_skip(7, 60_000, C.byref(used)) == 0
    assert used.value > 3_000          # ~8 % of the cases fast-forward (10^3 - 10^5 iterations each, executed here one by one)
