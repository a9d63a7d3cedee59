"""Parity tests proper: the CUDA path (through the C ABI of libkmeans_gpu.so) against the CPU
oracle on identical inputs, against the reference's committed golden images, and — at BASELINE
sizes — through size-independent properties.

Bar (north_star): labels / output pixels bit-exact, centroids within 1e-4 Lab.  What is asserted
here is stricter: everything is bit-exact against the oracle in its fixed-point sum mode
(centroids included), and within 2e-5 Lab of the oracle's f64-sum mode.
"""
import threading

import numpy as np
import pytest

from conftest import GOLDEN, load_rgba

pytestmark = pytest.mark.gpu

DARK_WHITE_RED = np.array([[5, 5, 5, 255], [255, 255, 255, 255], [255, 0, 0, 255]], np.uint8)


@pytest.fixture(scope="module")
def K():
    import kmeans_gpu_b200

    return kmeans_gpu_b200


@pytest.fixture(scope="module")
def D():
    import kmeans_gpu_b200.device as dev

    return dev


@pytest.fixture(scope="module")
def torch():
    import torch as t

    assert t.cuda.is_available(), "GPU tests need a CUDA device"
    return t


def dev_rgba(torch, arr):
    return torch.from_numpy(np.ascontiguousarray(arr)).cuda()


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def random_centroids(rng, k):
    c = np.empty((k, 4), np.float32)
    c[:, 0] = rng.uniform(0, 100, k)
    c[:, 1] = rng.uniform(-80, 90, k)
    c[:, 2] = rng.uniform(-100, 90, k)
    c[:, 3] = 1.0
    return c


# ---- K1/K3, table, K15 -------------------------------------------------------------------------

def test_srgb_table_bit_exact(proc, D, oracle):
    assert np.array_equal(bits(D.srgb_table(proc)), bits(oracle.srgb_decode_table() * np.float32(100.0)))


def test_convert_lab_bit_exact_random_and_grey(proc, D, oracle, torch):
    rng = np.random.default_rng(1)
    px = rng.integers(0, 256, (1 << 20, 4), dtype=np.uint8)
    px[:256, :3] = np.arange(256, dtype=np.uint8)[:, None]  # greys incl. both linear segments
    work = D.convert(proc, dev_rgba(torch, px)).cpu().numpy()
    want = oracle.convert(px)
    assert np.array_equal(bits(work[:, :3]), bits(want[:, :3]))
    # 4th float carries the exact chroma sqrt(a^2 + b^2)
    a, b = want[:, 1], want[:, 2]
    assert np.array_equal(bits(work[:, 3]), bits(np.sqrt(a * a + b * b, dtype=np.float32)))


def test_convert_lab_all_16m_colours(proc, D, oracle, torch):
    """H8: exhaustive over 2^24 colours.  Both sides round a double-precision pow to f32; a
    disagreement needs the double result within ~1e-16 relative of an f32 rounding boundary."""
    v = np.arange(1 << 24, dtype=np.uint32)
    px = np.stack([v & 255, (v >> 8) & 255, (v >> 16) & 255, np.full_like(v, 255)], axis=1).astype(np.uint8)
    work = D.convert(proc, dev_rgba(torch, px)).cpu().numpy()
    want = oracle.convert(px)
    diff = bits(work[:, :3]) != bits(want[:, :3])
    assert diff.sum() == 0, f"{diff.sum()} of {diff.size} components differ"


def test_convert_rgb_bit_exact(proc, D, K, oracle, torch):
    rng = np.random.default_rng(2)
    px = rng.integers(0, 256, (100003, 4), dtype=np.uint8)
    work = D.convert(proc, dev_rgba(torch, px), K.ColorSpace.Rgb).cpu().numpy()
    want = oracle.convert(px, oracle.RGB)
    assert np.array_equal(bits(work[:, :3]), bits(want[:, :3]))


@pytest.mark.parametrize("shape,max_dim", [((513, 768), 256), ((768, 513), 256), ((300, 301), 256), ((1080, 1920), 256),
                                           ((513, 768), 128), ((257, 3), 256), ((5, 1000), 256)])
def test_resize_bit_exact(proc, K, oracle, tokyo, shape, max_dim):
    h, w = shape
    if (h, w) == tokyo.shape[:2]:
        img = tokyo
    else:
        img = oracle.synth(w * h, seed=9, blobs=0).reshape(h, w, 4)
        img[..., 3] = np.arange(h * w, dtype=np.uint32).reshape(h, w) % 251  # alpha is filtered too
    out = proc.resize(img, max_dim)
    dw, dh = oracle.resized_dims(w, h, max_dim)
    assert out.dimensions == (dw, dh)
    assert np.array_equal(out.rgba, oracle.resize(img, dw, dh))


# ---- K5 ----------------------------------------------------------------------------------------

@pytest.mark.parametrize("k", [1, 2, 3, 8, 16, 17, 32, 33, 64, 256, 700])
def test_assign_bit_exact(proc, D, oracle, torch, k):
    rng = np.random.default_rng(100 + k)
    px = rng.integers(0, 256, (200000, 4), dtype=np.uint8)
    lab = oracle.convert(px)
    cent = random_centroids(rng, k)
    if k >= 3:
        cent[k - 1] = cent[0]  # exact duplicate: lowest index must win
        cent[1, :3] = lab[5, :3]  # a centroid sitting exactly on a pixel
    work = torch.from_numpy(np.concatenate([lab[:, :3], np.sqrt(lab[:, 1:2] ** 2 + lab[:, 2:3] ** 2, dtype=np.float32)], axis=1)).cuda()
    labels = D.assign(proc, work, cent).cpu().numpy().astype(np.uint32)
    want = oracle.assign(lab, cent)
    assert np.array_equal(labels, want)


def test_assign_adversarial_near_ties(proc, D, oracle, torch):
    """Centroids that differ by a few ulps, pixels equidistant from mirrored centroids: the fast
    score cannot separate them, so the exact path must reproduce the reference's strict-'<' scan."""
    rng = np.random.default_rng(5)
    base = random_centroids(rng, 8)
    cent = np.concatenate([base, base.copy(), base.copy()])
    cent[8:16, 0] = np.nextafter(cent[8:16, 0], np.float32(1000))  # +1 ulp in L
    cent[16:24, 1] = np.nextafter(cent[16:24, 1], np.float32(-1000))  # -1 ulp in a
    cent = cent[rng.permutation(24)]
    px = rng.integers(0, 256, (300000, 4), dtype=np.uint8)
    lab = oracle.convert(px)
    # mirrored pair around every 7th pixel in L
    lab2 = lab.copy()
    work_np = np.concatenate([lab2[:, :3], np.sqrt(lab2[:, 1:2] ** 2 + lab2[:, 2:3] ** 2, dtype=np.float32)], axis=1)
    labels = D.assign(proc, torch.from_numpy(work_np).cuda(), cent).cpu().numpy().astype(np.uint32)
    want, best, second = oracle.assign(lab2, cent, with_margin=True)
    assert np.array_equal(labels, want)
    assert ((second - best) == 0).mean() > 0.01  # the data really contains exact ties


# ---- K8-K11 ------------------------------------------------------------------------------------

def test_init_picks_tokyo(proc, D, K, oracle, torch, tokyo):
    sh = oracle.shrunk(tokyo)
    h, w = sh.shape[:2]
    work = D.convert(proc, dev_rgba(torch, sh))
    job = D.Job(proc, work, w, h, 8)
    idx, dist = job.init()
    lab = oracle.convert(sh)
    cent, oidx, odist = oracle.init(lab, w, h, 8, 144, 159)
    assert idx.tolist() == oidx.tolist()
    assert np.array_equal(bits(dist), bits(odist))
    assert np.array_equal(bits(job.centroids()), bits(cent))
    job.close()


@pytest.mark.parametrize("w,h,k", [(64, 64, 5), (100, 37, 12), (17, 3, 4), (256, 144, 16), (31, 1, 40)])
def test_init_picks_with_ties(proc, D, K, oracle, torch, w, h, k):
    """Few distinct colours => many exactly tied maxima and, once colours run out, zero maxima:
    exercises the selectCandidate tie rule (later 16-pixel chunk wins, earliest inside a chunk)."""
    rng = np.random.default_rng(w * h + k)
    pal = rng.integers(0, 256, (6, 4), dtype=np.uint8)
    pal[:, 3] = 255
    img = pal[rng.integers(0, 6, (h, w))]
    work = D.convert(proc, dev_rgba(torch, img))
    opts = K.Opts(seed_x=w // 3, seed_y=h // 2)
    job = D.Job(proc, work, w, h, k, opts=opts)
    idx, dist = job.init()
    cent, oidx, odist = oracle.init(oracle.convert(img), w, h, k, w // 3, h // 2)
    assert idx.tolist() == oidx.tolist()
    assert np.array_equal(bits(dist), bits(odist))
    assert np.array_equal(bits(job.centroids()), bits(cent))
    job.close()


# ---- K5 + K6/K7 fused pass, Lloyd loop -----------------------------------------------------------

@pytest.mark.parametrize("k,n_side", [(1, 64), (2, 300), (8, 512), (9, 200), (16, 400), (24, 333), (32, 256), (48, 300),
                                      (256, 256), (700, 160), (2100, 128)])
def test_lloyd_pass_bit_exact(proc, D, K, oracle, torch, k, n_side):
    w = h = n_side
    img = oracle.synth(w * h, seed=k, blobs=2 * k).reshape(h, w, 4)
    lab = oracle.convert(img)
    work = D.convert(proc, dev_rgba(torch, img))
    rng = np.random.default_rng(k)
    cent = lab[rng.choice(w * h, k, replace=False)].copy()
    cent[:, 3] = 1.0
    job = D.Job(proc, work, w, h, k)
    job.set_centroids(cent)
    ocent = cent
    for _ in range(3):
        job.step(1)
        labels = oracle.assign(lab, ocent)
        ocent, oconv, counts = oracle.update(lab, labels, ocent, 1.0, sum_mode=1)
        st = job.stats()
        assert np.array_equal(job.sums(), oracle.partial_sums(lab, labels, k))
        assert np.array_equal(bits(job.centroids()), bits(ocent))
        assert st["converged"] == oconv
    assert job.stats()["passes"] == 3
    job.close()


def test_block_accumulators_drain_mid_pass(K, D, oracle, torch, monkeypatch):
    """k > 32: block accumulators in shared memory are drained every 2^19 pixels per block in
    production; with the interval lowered to 2^11 every block drains many times in one pass."""
    monkeypatch.setenv("KMG_BLOCKACC_FLUSH_LOG2", "11")
    p = K.ImageProcessor(0)
    try:
        k, w, h = 40, 1400, 900
        img = oracle.synth(w * h, seed=9, blobs=80).reshape(h, w, 4)
        lab = oracle.convert(img)
        work = D.convert(p, dev_rgba(torch, img))
        cent = lab[np.random.default_rng(9).choice(w * h, k, replace=False)].copy()
        cent[:, 3] = 1.0
        job = D.Job(p, work, w, h, k)
        job.set_centroids(cent)
        job.step(1)
        labels = oracle.assign(lab, cent)
        assert np.array_equal(job.sums(), oracle.partial_sums(lab, labels, k))
        job.close()
    finally:
        p.close()


@pytest.mark.parametrize("k,env,n_variants", [(7, "KMG_LLOYD8_VARIANT", 13), (13, "KMG_LLOYD16_VARIANT", 4),
                                              (27, "KMG_LLOYD32_VARIANT", 2)])
def test_every_lloyd_variant_gives_the_same_sums(K, D, oracle, torch, monkeypatch, k, env, n_variants):
    """LLOYD_VARIANTS (kmg_api.cu): shared-memory or constant-bank table, atomic or read-modify-write
    slots, different geometries — every one must produce the oracle's integer sums."""
    w, h = 700, 500
    img = oracle.synth(w * h, seed=k, blobs=2 * k).reshape(h, w, 4)
    lab = oracle.convert(img)
    cent = lab[np.random.default_rng(k).choice(w * h, k, replace=False)].copy()
    cent[:, 3] = 1.0
    want = oracle.partial_sums(lab, oracle.assign(lab, cent), k)
    for v in range(n_variants):
        monkeypatch.setenv(env, str(v))
        p = K.ImageProcessor(0)
        try:
            work = D.convert(p, dev_rgba(torch, img))
            job = D.Job(p, work, w, h, k)
            job.set_centroids(cent)
            job.step(1)
            assert np.array_equal(job.sums(), want), (env, v)
            job.close()
        finally:
            p.close()


def test_more_live_jobs_than_constant_bank_slots(proc, D, K, oracle, torch):
    """64 slots of the constant bank per device: job 65 and later fall back to the shared-memory
    table; all 80 jobs alive at once must agree with the oracle."""
    w, h, k = 320, 240, 6
    img = oracle.synth(w * h, seed=3, blobs=12).reshape(h, w, 4)
    lab = oracle.convert(img)
    cent = lab[np.random.default_rng(3).choice(w * h, k, replace=False)].copy()
    cent[:, 3] = 1.0
    want = oracle.partial_sums(lab, oracle.assign(lab, cent), k)
    work = D.convert(proc, dev_rgba(torch, img))
    jobs = [D.Job(proc, work, w, h, k) for _ in range(80)]
    for j in jobs:
        j.set_centroids(cent)
        j.step(1)
    for j in jobs:
        assert np.array_equal(j.sums(), want)
        j.close()


def test_lloyd_empty_cluster_keeps_centroid(proc, D, K, oracle, torch):
    """choose_centroid.wgsl:185-194: an empty cluster keeps its centroid and never counts as
    converged, so the loop runs to the 128-iteration cap (core/src/modules.rs:764-766)."""
    w = h = 64
    img = oracle.synth(w * h, seed=3, blobs=4).reshape(h, w, 4)
    lab = oracle.convert(img)
    work = D.convert(proc, dev_rgba(torch, img))
    cent = np.array([[50, 0, 0, 1], [50, 0, 0, 1], [20, 10, 10, 1], [500, 500, 500, 1]], np.float32)
    job = D.Job(proc, work, w, h, 4)
    job.set_centroids(cent)
    passes = job.run()
    assert passes == 128
    c = job.centroids()
    assert np.array_equal(c[3], cent[3])  # never attracts a pixel, never moves
    # oracle agrees
    ocent = cent
    for _ in range(128):
        ocent, _, _ = oracle.update(lab, oracle.assign(lab, ocent), ocent, 1.0, sum_mode=1)
    assert np.array_equal(bits(c), bits(ocent))
    job.close()


def test_kmeans_tokyo_bit_exact(proc, K, oracle, tokyo):
    cent, passes = proc.kmeans_centroids(8, tokyo)
    ocent, opasses = oracle.kmeans(tokyo, 8, opts=oracle.default_opts(sum_mode=1))
    assert passes == opasses == 17
    assert np.array_equal(bits(cent), bits(ocent))
    # and within tolerance of the f64-sum restatement (north_star: 1e-4 Lab)
    fcent, _ = oracle.kmeans(tokyo, 8, opts=oracle.default_opts(sum_mode=0))
    assert np.abs(cent - fcent).max() < 2e-5


@pytest.mark.parametrize("k,cs", [(1, 0), (2, 0), (5, 0), (16, 0), (40, 0), (8, 1)])
def test_kmeans_various_k_bit_exact(proc, K, oracle, tokyo, k, cs):
    cent, passes = proc.kmeans_centroids(k, tokyo, K.ColorSpace(cs))
    ocent, opasses = oracle.kmeans(tokyo, k, cs, opts=oracle.default_opts(sum_mode=1))
    assert passes == opasses
    assert np.array_equal(bits(cent), bits(ocent))


def test_kmeans_no_shrink_and_options(proc, K, oracle):
    img = oracle.synth(500 * 300, seed=2, blobs=16).reshape(300, 500, 4)
    opts = K.Opts(max_dim=0, max_iter=20, check_every=4, seed_x=7, seed_y=9)
    cent, passes = proc.kmeans_centroids(8, img, opts=opts)
    ocent, opasses = oracle.kmeans(img, 8, opts=oracle.default_opts(sum_mode=1, max_dim=0, max_iter=20, check_every=4, seed_x=7, seed_y=9))
    assert passes == opasses
    assert np.array_equal(bits(cent), bits(ocent))


# ---- remap ---------------------------------------------------------------------------------------

def test_find_goldens_bit_exact(proc, K, tokyo):
    """The reference's own committed outputs (samples.sh:6-8)."""
    out = proc.find(tokyo, DARK_WHITE_RED, K.ReduceMode.Replace)
    assert np.array_equal(out.rgba, load_rgba("tokyo-find-replace-dark-white-red.png"))
    out = proc.find(tokyo, DARK_WHITE_RED, K.ReduceMode.Dither)
    assert np.array_equal(out.rgba, load_rgba("tokyo-find-dither-dark-white-red.png"))
    out = proc.find(tokyo, K.parse_palette(GOLDEN / "apollo-1x.png"), K.ReduceMode.Dither)
    assert np.array_equal(out.rgba, load_rgba("tokyo-find-dither-apollo.png"))


@pytest.mark.parametrize("mode", ["replace", "dither"])
@pytest.mark.parametrize("k", [1, 2, 7, 8, 16, 31, 64, 300])
def test_remap_bit_exact_synthetic(proc, K, oracle, mode, k):
    rng = np.random.default_rng(k)
    w, h = 517, 389  # odd width: dither groups straddle rows, tail group is partial
    img = oracle.synth(w * h, seed=k, blobs=0).reshape(h, w, 4)
    cols = rng.integers(0, 256, (k, 4), dtype=np.uint8)
    cols[:, 3] = 255
    cent = K.fixed_centroids(cols)
    m = K.ReduceMode.Replace if mode == "replace" else K.ReduceMode.Dither
    out = proc.remap(img, cent, m)
    want = (oracle.remap_replace if mode == "replace" else oracle.remap_dither)(img, cent)
    assert np.array_equal(out.rgba, want)


@pytest.mark.parametrize("mode", ["replace", "dither"])
def test_remap_near_grey_pixels_saturated_palette(proc, K, oracle, mode):
    """Worst case for the approximate-Lab certificate: near-grey pixels (hue undefined, the hue term of
    the distance has its largest gradient there) against pairs of saturated, nearly equidistant
    centroids, and far-away pixels whose nearest centroids are ~100 dE94 apart from them."""
    g = np.arange(256, dtype=np.uint8)
    rows = []
    for dr, dg, db in ((0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (2, 0, 1), (0, 3, 0), (1, 1, 0), (4, 0, 4)):
        rows.append(np.stack([np.clip(g.astype(int) + dr, 0, 255), np.clip(g.astype(int) + dg, 0, 255),
                              np.clip(g.astype(int) + db, 0, 255), np.full(256, 255)], axis=1))
    img = np.tile(np.stack(rows).astype(np.uint8), (16, 3, 1))  # 128 x 768
    pal = np.array([[255, 0, 0, 255], [0, 255, 0, 255], [0, 0, 255, 255], [255, 255, 0, 255], [255, 0, 255, 255],
                    [0, 255, 255, 255], [254, 0, 1, 255], [1, 254, 0, 255]], np.uint8)
    m = K.ReduceMode.Replace if mode == "replace" else K.ReduceMode.Dither
    f = oracle.remap_replace if mode == "replace" else oracle.remap_dither
    for cols in (pal, pal[:2], pal[[0, 6]], pal[[2, 5, 4]]):
        cent = K.fixed_centroids(cols)
        assert np.array_equal(proc.remap(img, cent, m).rgba, f(img, cent))


def test_remap_resurrect64_dither_4k(proc, K, oracle):
    """BASELINE config 3 at a size the oracle still finishes quickly (960x540 crop of the 4K case)."""
    pal = K.parse_palette(GOLDEN / "resurrect_64.png")
    img = oracle.synth(960 * 540, seed=1, blobs=0).reshape(540, 960, 4)
    out = proc.find(img, pal, K.ReduceMode.Dither)
    assert np.array_equal(out.rgba, oracle.find(img, pal, "dither"))


def test_remap_rgb_colour_space(proc, K, oracle):
    rng = np.random.default_rng(3)
    img = oracle.synth(300 * 200, seed=4, blobs=0).reshape(200, 300, 4)
    cent = K.fixed_centroids(rng.integers(0, 256, (9, 4), dtype=np.uint8), K.ColorSpace.Rgb)
    for m, f in ((K.ReduceMode.Replace, oracle.remap_replace), (K.ReduceMode.Dither, oracle.remap_dither)):
        out = proc.remap(img, cent, m, K.ColorSpace.Rgb)
        assert np.array_equal(out.rgba, f(img, cent, oracle.RGB))


@pytest.mark.parametrize("k", [1, 2, 3, 8, 46])
def test_meld_matches_oracle(proc, K, oracle, tokyo, k):
    """R10 (no golden exists): continuous output, exact arithmetic on both sides."""
    rng = np.random.default_rng(k)
    cols = rng.integers(0, 256, (k, 4), dtype=np.uint8)
    cols[:, 3] = 255
    img = tokyo[100:300, 200:500]
    cent = K.fixed_centroids(cols)
    out = proc.remap(img, cent, K.ReduceMode.Meld)
    want = oracle.remap_meld(img, cent)
    assert np.array_equal(out.rgba, want)


def test_reduce_tokyo_end_to_end(proc, K, oracle, tokyo):
    """BASELINE configs 1 and 2: reduce -c 8 (replace, dither) and palette -c 8."""
    for mode, name in ((K.ReduceMode.Replace, "replace"), (K.ReduceMode.Dither, "dither")):
        out, cent, passes = proc.reduce(8, tokyo, K.Algorithm.Kmeans, mode, return_details=True)
        want, ocent, opasses = oracle.reduce(tokyo, 8, name)
        assert passes == opasses == 17
        assert np.array_equal(bits(cent), bits(ocent))
        assert np.array_equal(out.rgba, want)
    pal = proc.palette(8, tokyo)
    assert np.array_equal(pal, oracle.palette(tokyo, 8))
    strip = load_rgba("tokyo-palette-c8-kmeans-s40.png")
    gold = np.array([strip[20, 20 + 40 * i] for i in range(8)]).astype(int)
    assert np.abs(pal[:, :3].astype(int) - gold[:, :3]).max() <= 1  # reference golden, +-1 LSB


def test_reduce_batch_matches_single(proc, K, oracle):
    frames = np.stack([oracle.synth(320 * 180, seed=3, blobs=32, frame=f).reshape(180, 320, 4) for f in range(3)])
    out, cent, passes = proc.reduce_batch(16, frames, K.ReduceMode.Dither)
    for f in range(3):
        want, ocent, opasses = oracle.reduce(frames[f], 16, "dither")
        assert passes[f] == opasses
        assert np.array_equal(bits(cent[f]), bits(ocent))
        assert np.array_equal(out[f], want)


# ---- fused (one thread-block cluster per image) vs staged k-means ---------------------------------

@pytest.mark.parametrize("w,h,k,cs", [(256, 171, 8, 0), (256, 144, 16, 0), (256, 256, 16, 0), (256, 256, 32, 0),
                                      (200, 256, 3, 0), (97, 61, 8, 0), (33, 7, 5, 0), (5, 3, 2, 0), (1, 1, 1, 0),
                                      (3, 1, 4, 0), (256, 200, 17, 0), (128, 128, 24, 1), (255, 255, 9, 1),
                                      (640, 360, 12, 0), (1000, 30, 6, 0)])
def test_fused_kmeans_matches_staged_and_oracle(proc, K, oracle, w, h, k, cs):
    """kmg_small.cuh against the stage-by-stage launches and the oracle: centroids, pass counts and
    the remapped image bit for bit (clusters of 16 CTAs for single images)."""
    img = oracle.synth(w * h, seed=11 + k, blobs=max(2, 2 * k)).reshape(h, w, 4)
    mode = K.ReduceMode.Dither if k % 2 == 0 else K.ReduceMode.Replace
    space = K.ColorSpace(cs)
    out_f, cent_f, passes_f = proc.reduce(k, img, reduce_mode=mode, color_space=space, return_details=True)
    out_s, cent_s, passes_s = proc.reduce(k, img, reduce_mode=mode, color_space=space, return_details=True,
                                          opts=K.Opts(fused_kmeans=False))
    assert passes_f == passes_s
    assert np.array_equal(bits(cent_f), bits(cent_s))
    assert np.array_equal(out_f.rgba, out_s.rgba)
    ocent, opasses = oracle.kmeans(img, k, cs, opts=oracle.default_opts(sum_mode=1))
    assert passes_f == opasses
    assert np.array_equal(bits(cent_f), bits(ocent))


def test_fused_kmeans_options_and_ties(proc, K, oracle):
    """Explicit seed, other iteration limits, no shrink; an image of two flat colours (every
    init distance tied, empty clusters, duplicate centroids)."""
    img = oracle.synth(240 * 100, seed=2, blobs=16).reshape(100, 240, 4)
    for kw in (dict(max_dim=0, max_iter=20, check_every=4, seed_x=7, seed_y=9), dict(max_dim=64, max_iter=3),
               dict(max_iter=1), dict(check_every=0, max_iter=11), dict(convergence=0.05)):
        cent, passes = proc.kmeans_centroids(8, img, opts=K.Opts(**kw))
        ocent, opasses = oracle.kmeans(img, 8, opts=oracle.default_opts(sum_mode=1, **kw))
        assert passes == opasses, kw
        assert np.array_equal(bits(cent), bits(ocent)), kw
    flat = np.zeros((64, 80, 4), np.uint8)
    flat[..., 3] = 255
    flat[:, 40:, :3] = (200, 30, 90)
    for k in (1, 2, 5, 16):
        out, cent, passes = proc.reduce(k, flat, return_details=True)
        want, ocent, opasses = oracle.reduce(flat, k, "replace")
        assert passes == opasses
        assert np.array_equal(bits(cent), bits(ocent))
        assert np.array_equal(out.rgba, want)


@pytest.mark.parametrize("k,mode", [(16, "dither"), (8, "replace"), (20, "dither"), (40, "replace")])
def test_reduce_batch_device_and_host_pipeline(proc, D, K, oracle, torch, k, mode):
    """BASELINE config 5 in miniature: frames resident in HBM (one cluster launch + one remap launch
    for the whole batch) and the chunked host pipeline, against per-frame reduce and the oracle."""
    n, w, h = 37, 480, 270
    frames = np.stack([oracle.synth(w * h, seed=3, blobs=32, frame=f).reshape(h, w, 4) for f in range(n)])
    rm = K.ReduceMode.Dither if mode == "dither" else K.ReduceMode.Replace
    dev_out, dev_cent, dev_passes = D.reduce_batch(proc, dev_rgba(torch, frames), k, rm)
    host_out, host_cent, host_passes = proc.reduce_batch(k, frames, rm)
    assert np.array_equal(dev_out.cpu().numpy(), host_out)
    assert np.array_equal(bits(dev_cent), bits(host_cent)) and np.array_equal(dev_passes, host_passes)
    for f in (0, 1, 17, n - 1):
        single, cent, passes = proc.reduce(k, frames[f], reduce_mode=rm, return_details=True)
        assert passes == host_passes[f]
        assert np.array_equal(bits(cent), bits(host_cent[f]))
        assert np.array_equal(single.rgba, host_out[f])
    want, ocent, opasses = oracle.reduce(frames[5], k, mode)
    assert opasses == host_passes[5]
    assert np.array_equal(bits(ocent), bits(host_cent[5]))
    assert np.array_equal(want, host_out[5])


@pytest.mark.parametrize("k,w,h,max_dim", [(16, 320, 180, 256), (8, 300, 200, 256), (24, 200, 256, 256), (5, 64, 48, 16)])
def test_reduce_batch_throughput_mode(proc, D, K, oracle, torch, k, w, h, max_dim):
    """Large batches run one persistent CTA per SM with the planes in an L2-resident scratch (>= 40
    frames): same results as the cluster launch of a single image and as the oracle."""
    n = 331  # more frames than SMs x resident CTAs: every CTA loops over several frames
    frames = np.stack([oracle.synth(w * h, seed=5, blobs=2 * k, frame=f).reshape(h, w, 4) for f in range(n)])
    opts = K.Opts(max_dim=max_dim)
    out, cent, passes = D.reduce_batch(proc, dev_rgba(torch, frames), k, K.ReduceMode.Dither, opts=opts)
    out = out.cpu().numpy()
    for f in (0, 1, 147, 148, 149, 295, 296, n - 1):
        single, c1, p1 = proc.reduce(k, frames[f], reduce_mode=K.ReduceMode.Dither, return_details=True, opts=opts)
        assert p1 == passes[f], f
        assert np.array_equal(bits(c1), bits(cent[f])), f
        assert np.array_equal(single.rgba, out[f]), f
    want, ocent, opasses = oracle.reduce(frames[200], k, "dither", opts=oracle.default_opts(sum_mode=1, max_dim=max_dim))
    assert opasses == passes[200]
    assert np.array_equal(bits(ocent), bits(cent[200]))
    assert np.array_equal(want, out[200])


def test_reduce_batch_many_chunks(proc, K, oracle):
    """More frames than one pipeline chunk holds (chunks of >= 16 frames through three workspaces)."""
    n, w, h = 70, 1920, 1080
    base = oracle.synth(w * h, seed=3, blobs=32, frame=0).reshape(h, w, 4)
    frames = np.empty((n, h, w, 4), np.uint8)
    for f in range(n):
        frames[f] = np.roll(base, 97 * f, axis=1)
        frames[f, :8, :8, :3] = f  # make every frame distinct
    out, cent, passes = proc.reduce_batch(16, frames, K.ReduceMode.Dither)
    for f in (0, 15, 16, 33, n - 1):
        single, c1, p1 = proc.reduce(16, frames[f], reduce_mode=K.ReduceMode.Dither, return_details=True)
        assert p1 == passes[f]
        assert np.array_equal(bits(c1), bits(cent[f]))
        assert np.array_equal(single.rgba, out[f])


# ---- boundary behaviour --------------------------------------------------------------------------

def test_errors(proc, K, tokyo):
    with pytest.raises(K.KmgError) as e:
        proc.reduce(0, tokyo)
    assert e.value.code == 1
    with pytest.raises(K.KmgError):
        proc.remap(tokyo, np.zeros((0, 4), np.float32))
    with pytest.raises(K.KmgError):
        proc.reduce(5000, tokyo)  # above MAX_K
    with pytest.raises(K.KmgError):
        proc.kmeans_centroids(4, tokyo, opts=K.Opts(seed_x=100000, seed_y=0))
    # the context is still usable afterwards
    assert proc.find(tokyo[:8, :8], DARK_WHITE_RED).dimensions == (8, 8)


def test_no_8192_cap_and_tiny_images(proc, K, oracle):
    """The reference is limited to 8192x8192 textures (README.md:9-11); linear buffers are not."""
    img = oracle.synth(9000 * 6, seed=1, blobs=4).reshape(6, 9000, 4)
    out = proc.reduce(3, img, reduce_mode=K.ReduceMode.Dither)
    want, _, _ = oracle.reduce(img, 3, "dither")
    assert np.array_equal(out.rgba, want)
    one = np.array([[[12, 200, 7, 255]]], np.uint8)
    out, cent, passes = proc.reduce(1, one, return_details=True)
    want, ocent, _ = oracle.reduce(one, 1, "replace")
    assert np.array_equal(out.rgba, want) and np.array_equal(bits(cent), bits(ocent))


def test_concurrent_callers(proc, K, oracle, tokyo):
    """core/examples/parallel.rs:36-51 — 14 threads, k = 2..15, one shared processor."""
    results = {}

    def work(k):
        results[k] = proc.reduce(k, tokyo, reduce_mode=K.ReduceMode.Replace, return_details=True)

    threads = [threading.Thread(target=work, args=(k,)) for k in range(2, 16)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for k in (2, 7, 15):
        want, ocent, opasses = oracle.reduce(tokyo, k, "replace")
        out, cent, passes = results[k]
        assert passes == opasses and np.array_equal(bits(cent), bits(ocent)) and np.array_equal(out.rgba, want)


def test_remap_of_a_large_host_image_runs_as_row_bands(proc, K, oracle):
    """kmg_remap pipelines host images >= 12 MiB as row bands over three streams (upload / kernel /
    read-back overlap).  Ragged geometry: width not a multiple of 4, rows not a multiple of the
    band height; the dither matrix must stay aligned across band boundaries."""
    w, h = 2021, 1667  # 13.5 MB -> four bands of 416 rows + a short one
    img = oracle.synth(w * h, seed=11, blobs=6).reshape(h, w, 4)
    pal = np.array([[12, 16, 20, 255], [200, 40, 30, 255], [240, 240, 230, 255], [30, 120, 200, 255], [90, 160, 60, 255]],
                   np.uint8)
    for mode, name in ((K.ReduceMode.Replace, "replace"), (K.ReduceMode.Dither, "dither"), (K.ReduceMode.Meld, "meld")):
        got = proc.find(img, pal, mode)
        want = oracle.find(img, pal, name)
        assert np.array_equal(got.rgba, want), name
    # pinned buffers (the overlapping case) give the same bytes
    pin_in = K.pinned_empty((h, w, 4))
    pin_in[...] = img
    pin_out = K.pinned_empty((h, w, 4))
    got = proc.find(pin_in, pal, K.ReduceMode.Dither, out=pin_out)
    assert np.array_equal(pin_out, oracle.find(img, pal, "dither"))


def test_banded_upload_of_a_large_unshrunk_image(proc, K, oracle):
    """Host images >= 32 MiB that are clustered at full size are uploaded in bands on a second
    stream and converted band by band (the conversion hides behind the upload)."""
    w, h = 3001, 2803  # 33.6 MB, ragged against the 16 MiB bands
    img = oracle.synth(w * h, seed=21, blobs=8).reshape(h, w, 4)
    opts = K.Opts(max_dim=0, max_iter=3, check_every=0)
    cent, passes = proc.kmeans_centroids(4, img, opts=opts)
    ocent, opasses = oracle.kmeans(img, 4, oracle.LAB, oracle.default_opts(max_dim=0, max_iter=3, check_every=0))
    assert passes == opasses and np.array_equal(bits(cent), bits(ocent))
    out = proc.reduce(4, img, reduce_mode=K.ReduceMode.Replace, opts=opts)
    assert np.array_equal(out.rgba, oracle.remap_replace(img, ocent))


def test_reduce_of_a_large_host_image_is_pipelined(proc, K, oracle):
    """kmg_reduce on a host image >= 32 MiB that is shrunk before clustering: the rows the bilinear
    taps read are uploaded first and clustered, then the image streams through upload / remap /
    read-back bands.  Same bytes as the oracle's reduce; ragged geometry; pageable and pinned."""
    w, h = 3001, 2803  # 33.6 MB, bands of 1396 rows + a short one; shrinks to 256 x 239
    img = oracle.synth(w * h, seed=31, blobs=10).reshape(h, w, 4)
    n0 = proc.launch_count()
    out, cent, passes = proc.reduce(5, img, reduce_mode=K.ReduceMode.Dither, return_details=True)
    assert proc.launch_count() - n0 == 1 + 3  # one-launch k-means + one remap launch per band
    want, ocent, opasses = oracle.reduce(img, 5, "dither")
    assert passes == opasses and np.array_equal(bits(cent), bits(ocent))
    assert np.array_equal(out.rgba, want)
    pin_in, pin_out = K.pinned_empty((h, w, 4)), K.pinned_empty((h, w, 4))
    pin_in[...] = img
    proc.reduce(5, pin_in, reduce_mode=K.ReduceMode.Replace, out=pin_out)
    assert np.array_equal(pin_out, oracle.reduce(img, 5, "replace")[0])
    # portrait image (the shrink rule takes the other branch), staged k-means (no fused kernel)
    imgp = np.ascontiguousarray(img.transpose(1, 0, 2))
    outp = proc.reduce(5, imgp, reduce_mode=K.ReduceMode.Dither, opts=K.Opts(fused_kmeans=False))
    assert np.array_equal(outp.rgba, oracle.reduce(imgp, 5, "dither")[0])


def test_concurrent_staged_jobs_share_the_constant_bank(proc, K, oracle, tokyo):
    """Staged launches (no shrink, > 65,536 clustered pixels) with k <= 16 keep their table in a
    per-job slot of the constant bank: jobs running at the same time on different streams must not
    see each other's tables, and a reduce() (whose remap copies the job) must hand its slot back
    exactly once.  20 threads x 3 rounds; every result against the sequential call, one against
    the oracle."""
    img = np.ascontiguousarray(tokyo[:480, :640])  # 307,200 px: beyond what the one-launch k-means holds
    opts = K.Opts(max_dim=0, max_iter=12)
    n0 = proc.launch_count()
    proc.kmeans_centroids(8, img, opts=K.Opts(max_dim=0, max_iter=2))
    assert proc.launch_count() - n0 >= 10  # really the staged launches (convert, seed, 7 rounds, prepare, passes)
    ks = [3, 5, 8, 8, 9, 12, 16, 16, 7, 4] * 2
    seq = {k: proc.reduce(k, img, reduce_mode=K.ReduceMode.Dither, opts=opts, return_details=True) for k in set(ks)}
    errors = []

    def work(i, k):
        for _ in range(3):
            out, cent, passes = proc.reduce(k, img, reduce_mode=K.ReduceMode.Dither, opts=opts, return_details=True)
            sout, scent, spasses = seq[k]
            if passes != spasses or not np.array_equal(bits(cent), bits(scent)) or not np.array_equal(out.rgba, sout.rgba):
                errors.append((i, k))

    threads = [threading.Thread(target=work, args=(i, k)) for i, k in enumerate(ks)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    want, ocent, opasses = oracle.reduce(img, 8, "dither", oracle.default_opts(max_dim=0, max_iter=12))
    out, cent, passes = seq[8]
    assert passes == opasses and np.array_equal(bits(cent), bits(ocent)) and np.array_equal(out.rgba, want)
    # far more sequential jobs than slots (64): slots must come back
    for _ in range(80):
        proc.kmeans_centroids(6, img, opts=K.Opts(max_dim=0, max_iter=2))


# ---- the certificate --------------------------------------------------------------------------------

def test_fast_lab_error_bound(proc, D):
    err = D.fast_lab_error(proc)
    # kmg::fast::LAB_ERR = 1.5e-4 budgets 1.0e-4 for this term (exhaustive over all 2^24 colours)
    assert 0 < err < 1.0e-4, err
    print("max |fast Lab - exact Lab| =", err)


def test_exact_path_is_rare_on_natural_data(proc, D, K, oracle, torch, tokyo):
    sh = oracle.shrunk(tokyo)
    h, w = sh.shape[:2]
    work = D.convert(proc, dev_rgba(torch, sh))
    job = D.Job(proc, work, w, h, 8)
    job.init()
    job.step(10)
    st = job.stats()
    assert st["slow_pixels"] < 0.01 * 10 * w * h, st
    job.close()


# ---- full BASELINE sizes: size-independent properties ---------------------------------------------

def test_full_size_properties_8192(proc, D, K, oracle, torch):
    """BASELINE config 4 geometry (8192x8192, blobs(512), k=256) on the device generator."""
    w = h = 8192
    n = w * h
    img = D.synth(proc, n, seed=2, blobs=512).view(h, w, 4)
    # generator agrees with the oracle on a slice
    assert np.array_equal(img.view(-1, 4)[5_000_000:5_004_096].cpu().numpy(), oracle.synth(4096, first_pixel=5_000_000, seed=2, blobs=512))
    work = D.convert(proc, img)
    k = 256
    rows = torch.randint(0, n, (k,), generator=torch.Generator().manual_seed(1))
    cent = work[rows.cuda()].cpu().numpy().copy()
    cent[:, 3] = 1.0
    job = D.Job(proc, work, w, h, k, opts=K.Opts(max_dim=0))
    job.set_centroids(cent)
    job.step(1)
    c1 = job.centroids()
    # parity on a crop through the oracle: labels of 64k pixels and their contribution
    crop = slice(12_345_678, 12_345_678 + 65536)
    lab_crop = work[crop].cpu().numpy()
    lab_crop4 = lab_crop.copy()
    labels_gpu = D.assign(proc, work[crop].contiguous(), cent).cpu().numpy().astype(np.uint32)
    assert np.array_equal(labels_gpu, oracle.assign(lab_crop4, cent))
    # conservation: every pixel is counted exactly once
    s_all = job.sums()
    assert int(s_all[:, 3].sum()) == n
    assert np.isfinite(c1).all() and (c1[:, 0] >= 0).all() and (c1[:, 0] <= 100.001).all()
    # exactness under sharding: integer sums of two half-image jobs add up to the whole-image sums
    half = n // 2
    j0 = D.Job(proc, work[:half], w, h // 2, k, opts=K.Opts(max_dim=0))
    j1 = D.Job(proc, work[half:], w, h // 2, k, opts=K.Opts(max_dim=0))
    for j in (j0, j1):
        j.set_centroids(cent)
        j.step(1)
    assert np.array_equal(j0.sums() + j1.sums(), s_all)
    # and the finalisation of those sums is the oracle's
    ocent, _ = oracle.finalize(s_all, cent, 1.0)
    assert np.array_equal(bits(c1), bits(ocent))
    # replace remap at full size only emits palette colours
    out1 = job.remap(img, K.ReduceMode.Replace)
    pal = oracle.revert(c1)
    got = out1.view(torch.int32).unique().cpu().numpy().view(np.uint32)
    assert set(got.tolist()) <= set(pal.view(np.uint32).ravel().tolist())
    for j in (job, j0, j1):
        j.close()


# ---- certificate audit: every certified label is the reference scan's, on ALL 2^24 colours -----------

def _audit_palettes(K, rng, k, n):
    """n palettes of k centroids (Lab): random in-gamut colours with sub-LSB jitter, near-duplicates,
    saturated primaries, near-greys — the shapes that stress the error bound."""
    out = []
    for i in range(n):
        kind = i % 4
        cols = rng.integers(0, 256, (k, 4), dtype=np.uint8)
        if kind == 2:  # saturated: every channel 0 or 255 (plus a few mid values so that k colours exist)
            cols[:, :3] = np.where(rng.random((k, 3)) < 0.8, rng.integers(0, 2, (k, 3)) * 255, cols[:, :3])
        if kind == 3:  # near-grey: r ~ g ~ b within one or two levels
            g = rng.integers(0, 256, (k, 1))
            cols[:, :3] = np.clip(g + rng.integers(-2, 3, (k, 3)), 0, 255)
        cols[:, 3] = 255
        cent = K.fixed_centroids(cols, K.ColorSpace.Lab)
        if kind != 2:  # k-means centroids are means, not 8-bit colours
            cent[:, :3] += rng.normal(0, 0.2, (k, 3)).astype(np.float32)
        if kind == 1 and k >= 2:  # near-duplicates: half of the palette sits 1e-4 .. 1e-2 next to the other half
            h = k // 2
            cent[h:2 * h, :3] = cent[:h, :3] + rng.choice([1e-4, 1e-3, 1e-2], (h, 1)).astype(np.float32) * \
                rng.choice([-1.0, 1.0], (h, 3)).astype(np.float32)
        cent[:, 3] = 1.0
        out.append(cent)
    return out


def test_certificate_audit_all_colours(proc, D, K, torch):
    """The production searches trust a certified label without evaluating the reference distance.
    On all 2^24 sRGB colours x 204 palettes (k = 2 ... 700) every search (8- and 16-entry tables,
    chunked, the resident-table search of the k <= 8 Lloyd pass) in every role (Lloyd on the exact
    plane, remap replace, remap dither on the fast Lab) must certify only labels equal to the
    reference's in-order scan (find_centroid.wgsl:29-41).  kmg_dev_audit counts the exceptions."""
    v = torch.arange(1 << 24, dtype=torch.int32, device="cuda")
    img = (v | (255 << 24)).view(torch.uint8).view(4096, 4096, 4)
    work = D.convert(proc, img)
    rng = np.random.default_rng(2024)
    total = {"audits": 0, "uncertified": 0}
    plan = {2: 40, 8: 52, 16: 40, 64: 32, 256: 25, 700: 15}
    assert sum(plan.values()) >= 200
    for k, n_pal in plan.items():
        searches = ([0, 3] if k <= 8 else []) + ([1] if k <= 16 else []) + [2]
        for pal_i, cent in enumerate(_audit_palettes(K, rng, k, n_pal)):
            for mode in (0, 1, 2):
                for search in searches:
                    if search == 3 and mode != 0:
                        continue  # the resident-table search only exists in the Lloyd pass
                    if mode == 0:
                        wrong, unc = D.audit(proc, cent, search, 0, work=work, w=4096, h=4096)
                    else:
                        wrong, unc = D.audit(proc, cent, search, mode, rgba=img, w=4096, h=4096)
                    assert wrong == 0, (k, mode, search, wrong, cent.tolist())
                    if pal_i % 4 == 0:  # random palettes: the fast path does the work (near-duplicate palettes
                        assert unc < (0.05 if k <= 64 else 0.2) * (1 << 24), (k, mode, search, unc)  # legitimately send everything to the exact path)
                    total["audits"] += 1
                    total["uncertified"] += unc
    assert total["audits"] >= 600


def test_certificate_audit_rgb_colour_space(proc, D, K, torch):
    v = torch.arange(1 << 24, dtype=torch.int32, device="cuda")
    img = (v | (255 << 24)).view(torch.uint8).view(4096, 4096, 4)
    work = D.convert(proc, img, K.ColorSpace.Rgb)
    rng = np.random.default_rng(5)
    for k in (2, 8, 16, 40):
        for _ in range(4):
            cent = np.ones((k, 4), np.float32)
            cent[:, :3] = rng.random((k, 3), dtype=np.float32)
            searches = ([0, 3] if k <= 8 else []) + ([1] if k <= 16 else []) + [2]
            for search in searches:
                assert D.audit(proc, cent, search, 0, work=work, w=4096, h=4096, color_space=K.ColorSpace.Rgb)[0] == 0
                if search != 3:
                    for mode in (1, 2):
                        assert D.audit(proc, cent, search, mode, rgba=img, w=4096, h=4096, color_space=K.ColorSpace.Rgb)[0] == 0


# ---- lazy farthest-point rounds (kmg_init_lazy.cuh) against the one-sweep-per-round kernels -----------

@pytest.mark.parametrize("side,k,blobs,eager_rounds", [(1024, 64, 128, 8), (1536, 256, 512, 8), (700, 33, 0, 1), (512, 300, 4, 3),
                                                       (640, 12, 24, 2)])
def test_lazy_init_matches_full_sweeps(K, D, oracle, torch, monkeypatch, side, k, blobs, eager_rounds):
    """Same picks, same max-min distances (bit for bit), same centroids: the lazy rounds only skip
    pixels that provably cannot win.  blobs=4 with k=300: far more clusters than colours groups, the
    distances collapse and the threshold has to chase them down."""
    img = oracle.synth(side * side, seed=11, blobs=blobs).reshape(side, side, 4)
    res = []
    monkeypatch.setenv("KMG_INIT_LAZY_MIN_K", "0")
    monkeypatch.setenv("KMG_INIT_EAGER_ROUNDS", str(eager_rounds))
    for eager in ("1", "0"):
        monkeypatch.setenv("KMG_INIT_EAGER", eager)
        p = K.ImageProcessor(0)
        try:
            work = D.convert(p, dev_rgba(torch, img))
            job = D.Job(p, work, side, side, k, opts=K.Opts(max_dim=0))
            idx, dist = job.init()
            st = job.init_stats()
            res.append((idx.copy(), dist.copy(), job.centroids().copy(), st))
            job.close()
        finally:
            p.close()
    (i0, d0, c0, _), (i1, d1, c1, st) = res
    assert i0.tolist() == i1.tolist()
    assert np.array_equal(bits(d0), bits(d1)) and np.array_equal(bits(c0), bits(c1))
    assert st["sweeps"] >= k - 1 - eager_rounds and st["exact"] <= st["pairs"], st
    if k >= 64:
        assert st["refreshed"] < 0.25 * (k - 1 - eager_rounds) * side * side, st  # full sweeps would refresh every pixel every round


def test_lazy_init_flat_image_and_zero_maximum(K, D, oracle, torch, monkeypatch):
    """Two colours, k = 6: from round 3 on every distance is zero — the threshold must fall to 'every
    pixel' and the zero maximum must select pixel 0 (plus_plus_init.wgsl: Candidate(0, 0.0))."""
    w, h = 300, 200
    img = np.zeros((h, w, 4), np.uint8)
    img[..., 3] = 255
    img[:, : w // 3, 0] = 200
    monkeypatch.setenv("KMG_INIT_LAZY_MIN_K", "0")
    monkeypatch.setenv("KMG_INIT_EAGER_ROUNDS", "1")
    proc = K.ImageProcessor(0)
    try:
        work = D.convert(proc, dev_rgba(torch, img))
        job = D.Job(proc, work, w, h, 6, opts=K.Opts(max_dim=0))
        idx, dist = job.init()
        assert job.init_stats()["sweeps"] >= 4
        lab = oracle.convert(img)
        ocent, oidx, odist = oracle.init(lab, w, h, 6, int(w * 0.5625), int(h * 0.93359375))
        assert idx.tolist() == oidx.tolist() and np.array_equal(bits(dist), bits(odist))
        assert np.array_equal(bits(job.centroids()), bits(ocent))
        job.close()
    finally:
        proc.close()
