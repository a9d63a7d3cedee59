"""`Algorithm::Octree` (core/src/octree.rs, operations.rs:90-97, lib.rs:288-331): the host-side
quantiser kmg_octree_palette against the literal Python restatement oracle/octree_oracle.py, the
reference's own unit test as known answer, hand-checked small cases; on a GPU, the full
reduce/palette path (device shrink to <= 128 px, CPU quantiser, sort by L, fixed-palette remap)."""
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, str(ROOT / "oracle"))
import octree_oracle  # noqa: E402

# the 46 colours of the reference's test (core/src/octree.rs:285-332) — the apollo palette
REF_TEST_COLORS = [
    (9, 10, 20), (16, 20, 31), (21, 29, 40), (23, 32, 56), (25, 51, 45), (30, 29, 57), (32, 46, 55), (36, 21, 39),
    (37, 58, 94), (37, 86, 46), (52, 28, 39), (57, 74, 80), (60, 94, 139), (64, 39, 81), (65, 29, 49), (70, 130, 50),
    (77, 43, 50), (79, 143, 186), (87, 114, 119), (96, 44, 44), (115, 190, 211), (117, 36, 56), (117, 167, 67),
    (122, 54, 123), (122, 72, 65), (129, 151, 150), (136, 75, 43), (162, 62, 140), (164, 221, 219), (165, 48, 48),
    (168, 181, 178), (168, 202, 88), (173, 119, 87), (190, 119, 43), (192, 148, 115), (198, 81, 151), (199, 207, 204),
    (207, 87, 60), (208, 218, 145), (215, 181, 148), (218, 134, 62), (222, 158, 65), (223, 132, 165), (231, 213, 179),
    (232, 193, 112), (235, 237, 233),
]


def _rgba(colors):
    a = np.array([(r, g, b, 255) for r, g, b in colors], np.uint8)
    return a


@pytest.fixture(scope="module")
def K():
    import kmeans_gpu_b200

    return kmeans_gpu_b200


@pytest.fixture(scope="module")
def octree_colors(native_lib):
    from kmeans_gpu_b200.processor import octree_colors as f

    return f


def _as_tuples(a):
    return [tuple(int(v) for v in c) for c in np.asarray(a).reshape(-1, 4)]


def test_reference_unit_test_known_answer(octree_colors):
    # octree.rs:280-340: 46 colours reduce to exactly 8
    pix = _rgba(REF_TEST_COLORS)
    pal = octree_colors(pix, 8)
    assert len(pal) == 8
    assert _as_tuples(pal) == octree_oracle.octree_palette(pix, 8)
    assert _as_tuples(pal) == sorted(_as_tuples(pal))  # palette.sort()
    assert all(c[3] == 255 for c in _as_tuples(pal))


def test_hand_checked_cases(octree_colors):
    # fewer distinct colours than requested: every colour survives, sorted as tuples, de-duplicated
    pix = _rgba([(200, 10, 10), (10, 200, 10), (200, 10, 10), (10, 10, 200)])
    assert _as_tuples(octree_colors(pix, 8)) == [(10, 10, 200, 255), (10, 200, 10, 255), (200, 10, 10, 255)]
    # two leaves under the same depth-8 parent (they differ in the last bit of blue only) merge into
    # their integer mean before an unrelated colour is touched: (10,10,10) x1 + (10,10,11) x2 -> b = 32 // 3
    pix = _rgba([(10, 10, 10), (10, 10, 11), (10, 10, 11), (250, 250, 250)])
    assert _as_tuples(octree_colors(pix, 2)) == [(10, 10, 10, 255), (250, 250, 250, 255)]
    assert octree_oracle.octree_palette(pix, 2) == [(10, 10, 10, 255), (250, 250, 250, 255)]
    # color_count 0 -> empty (octree.rs:67-69); one colour
    assert len(octree_colors(pix, 0)) == 0
    assert _as_tuples(octree_colors(_rgba([(1, 2, 3)] * 5), 4)) == [(1, 2, 3, 255)]
    # empty input
    assert len(octree_colors(np.zeros((0, 4), np.uint8), 4)) == 0


@pytest.mark.parametrize("seed,n,k", [(0, 300, 1), (1, 300, 2), (2, 1000, 8), (3, 4096, 16), (4, 4096, 64), (5, 2000, 256),
                                      (6, 50, 3), (7, 16384, 5)])
def test_matches_literal_restatement(octree_colors, seed, n, k):
    rng = np.random.default_rng(seed)
    if seed % 2:
        # clustered colours: many duplicates and deep shared prefixes
        centres = rng.integers(0, 256, (12, 3))
        px = np.clip(centres[rng.integers(0, 12, n)] + rng.integers(-6, 7, (n, 3)), 0, 255)
    else:
        px = rng.integers(0, 256, (n, 3))
    pix = np.concatenate([px, np.full((n, 1), 255)], axis=1).astype(np.uint8)
    got = _as_tuples(octree_colors(pix, k))
    assert got == octree_oracle.octree_palette(pix, k)
    assert len(got) <= k


def test_tokyo_shrink_known_counts(octree_colors, tokyo, oracle):
    # the pixels octree_palette really sees: the <= 128 px shrink of the image (lib.rs:293-316)
    small = oracle.resize(tokyo, *oracle.resized_dims(tokyo.shape[1], tokyo.shape[0], 128))
    assert small.shape[:2] == (85, 128)
    for k in (2, 8, 16):
        got = _as_tuples(octree_colors(small, k))
        assert got == octree_oracle.octree_palette(small.reshape(-1, 4), k)
        assert 1 <= len(got) <= k


@pytest.mark.gpu
def test_reduce_and_palette_octree_end_to_end(proc, K, tokyo, oracle):
    small = oracle.resize(tokyo, *oracle.resized_dims(tokyo.shape[1], tokyo.shape[0], 128))
    want_pal = np.array(octree_oracle.octree_palette(small.reshape(-1, 4), 8), np.uint8)
    # lib.rs:318-329: sorted by the Lab L of the 8-bit colour
    L = oracle.pal_srgb8_to_lab(want_pal)[:, 0]
    want_pal = want_pal[np.argsort(L, kind="stable")]
    got_pal = proc.palette(8, tokyo, K.Algorithm.Octree)
    assert np.array_equal(got_pal, want_pal)
    for mode, name in ((K.ReduceMode.Replace, "replace"), (K.ReduceMode.Dither, "dither")):
        got = proc.reduce(8, tokyo, K.Algorithm.Octree, mode)
        want = oracle.find(tokyo, want_pal, name)
        assert np.array_equal(got.rgba, want)
    # small images are quantised as they are (no shrink below 128 px)
    tile = np.ascontiguousarray(tokyo[100:164, 200:300])
    got_pal = proc.palette(4, tile, K.Algorithm.Octree)
    want = np.array(octree_oracle.octree_palette(tile.reshape(-1, 4), 4), np.uint8)
    want = want[np.argsort(oracle.pal_srgb8_to_lab(want)[:, 0], kind="stable")]
    assert np.array_equal(got_pal, want)
