"""Pins the CPU oracle against every golden vector / known answer the reference holds for the
hot path: shader-test KATs and the committed sample images."""
import numpy as np
import pytest

from conftest import load_rgba

DARK_WHITE_RED = np.array([[5, 5, 5, 255], [255, 255, 255, 255], [255, 0, 0, 255]], np.uint8)


def hexes(cols):
    return ["#%02X%02X%02X" % tuple(int(v) for v in c[:3]) for c in cols]


def test_cie94_kat(oracle):
    # core/src/shader_tests.rs:180-186 — expects 19.094658 +- 0.01
    lab = oracle.pal_srgb8_to_lab(np.array([[255, 0, 0, 255], [255, 128, 0, 255]], np.uint8))
    d = oracle.cie94(lab[0, :3], lab[1, :3])
    assert abs(d - 19.094658) < 0.01
    # asymmetry (SC, SH come from the first argument)
    assert abs(oracle.cie94(lab[1, :3], lab[0, :3]) - 20.3009) < 0.01


def test_pow_kat(oracle):
    # core/src/shader_tests.rs:231-240 — 2.1^7 = 180.1088541 +- 0.1
    assert abs(oracle.pow_f32(2.1, 7.0) - 180.1088541) < 0.1


def test_find_replace_golden_bit_exact(oracle, tokyo):
    # samples.sh:6
    out = oracle.find(tokyo, DARK_WHITE_RED, "replace")
    assert np.array_equal(out, load_rgba("tokyo-find-replace-dark-white-red.png"))
    counts = [(out[..., :3] == c[:3]).all(axis=2).sum() for c in DARK_WHITE_RED]
    assert counts == [334541, 40844, 18599]


def test_find_dither_golden_bit_exact(oracle, tokyo):
    # samples.sh:7
    out = oracle.find(tokyo, DARK_WHITE_RED, "dither")
    assert np.array_equal(out, load_rgba("tokyo-find-dither-dark-white-red.png"))
    thr = oracle.dither_threshold(oracle.pal_srgb8_to_lab(DARK_WHITE_RED))
    assert abs(thr - 56.943558) < 1e-4


def test_find_dither_apollo_golden_bit_exact(oracle, tokyo):
    # samples.sh:8 — palette image sorted as RGBA tuples (cli/src/args.rs:208-210)
    import kmeans_gpu_b200 as K
    from conftest import GOLDEN

    pal = K.parse_palette(GOLDEN / "apollo-1x.png")
    assert pal.shape == (46, 4)
    out = oracle.find(tokyo, pal, "dither")
    assert np.array_equal(out, load_rgba("tokyo-find-dither-apollo.png"))


def test_palette_roundtrip_of_fixed_palettes(oracle):
    import kmeans_gpu_b200 as K
    from conftest import GOLDEN

    pal = np.concatenate([DARK_WHITE_RED, K.parse_palette(GOLDEN / "apollo-1x.png"), K.parse_palette(GOLDEN / "resurrect_64.png")])
    lab = oracle.pal_srgb8_to_lab(pal)
    assert np.array_equal(oracle.pal_lab_to_srgb8(lab), pal)
    # the shader-side reversion (K2) agrees on these colours too
    assert np.array_equal(oracle.revert(lab), pal)


def test_kmeans_intermediate_kats(oracle, tokyo):
    """Seed pixel, farthest-point picks, pass count and final centroids."""
    sh = oracle.shrunk(tokyo)
    assert sh.shape == (171, 256, 4)
    assert oracle.seed_pixel(256, 171) == (144, 159)
    lab = oracle.convert(sh)
    seed = lab[159 * 256 + 144]
    assert tuple(sh[159, 144, :3]) == (104, 13, 16)
    assert np.allclose(seed[:3], [21.0916, 38.6204, 24.8076], atol=2e-4)
    cent, idx, dist = oracle.init(lab, 256, 171, 8, 144, 159)
    picks = [(int(i % 256), int(i // 256)) for i in idx]
    assert picks == [(144, 159), (166, 122), (35, 79), (6, 36), (148, 69), (86, 120), (175, 88), (156, 67)]
    assert np.allclose(dist[1:], [91.2856, 54.9571, 50.4015, 42.5389, 38.8351, 35.1938, 35.1306], atol=2e-4)
    for sum_mode in (0, 1):
        c, passes = oracle.kmeans(tokyo, 8, opts=oracle.default_opts(sum_mode=sum_mode))
        assert passes == 17  # converges at the check of iteration 16
        want = np.array([[24.6338, 23.0883, 20.9086], [89.3919, -2.4502, 4.8480], [34.6962, -3.6075, 7.3194],
                         [5.0142, 0.2285, 1.5154], [39.5108, 51.6380, 41.6081], [58.1388, 30.4721, 35.1019],
                         [14.6309, 4.6639, 5.3817], [61.2368, -7.4443, 1.1875]], np.float32)
        assert np.allclose(c[:, :3], want, atol=2e-4)


def test_sum_modes_agree(oracle, tokyo):
    """Fixed-point (2^-15) sums and f64 sums give the same centroids to well below 1e-4 Lab."""
    c0, p0 = oracle.kmeans(tokyo, 8, opts=oracle.default_opts(sum_mode=0))
    c1, p1 = oracle.kmeans(tokyo, 8, opts=oracle.default_opts(sum_mode=1))
    assert p0 == p1
    assert np.abs(c0 - c1).max() < 2e-5


def _labels_from_image(img, palette_rgb):
    """Map each pixel to the index of its (exact) palette colour; -1 if absent."""
    key = img[..., 0].astype(np.int64) << 16 | img[..., 1].astype(np.int64) << 8 | img[..., 2].astype(np.int64)
    out = np.full(key.shape, -1, np.int64)
    for i, c in enumerate(palette_rgb):
        out[key == (int(c[0]) << 16 | int(c[1]) << 8 | int(c[2]))] = i
    return out


@pytest.mark.parametrize("mode,golden", [("replace", "tokyo-reduce-c8-kmeans-replace.png"),
                                         ("dither", "tokyo-reduce-c8-kmeans-dither.png")])
def test_reduce_goldens_within_one_lsb(oracle, tokyo, mode, golden):
    """samples.sh:3-4 — approximate pins: the author's GPU centroids differ in the 4th digit, so
    palette entries agree to +-1/255 and < 0.2 % of labels differ."""
    out, cent, passes = oracle.reduce(tokyo, 8, mode)
    gold = load_rgba(golden)
    ours = oracle.revert(cent)[:, :3].astype(int)
    gold_cols = np.unique(gold.reshape(-1, 4)[:, :3], axis=0).astype(int)
    assert len(gold_cols) == 8
    # match every golden colour to one of ours within +-1 per channel
    match = []
    for g in gold_cols:
        d = np.abs(ours - g).max(axis=1)
        assert d.min() <= 1, (g, ours)
        match.append(int(d.argmin()))
    assert sorted(match) == list(range(8))
    gl = _labels_from_image(gold, gold_cols)
    ol = _labels_from_image(out, ours)
    remap = np.array(match)
    mism = (remap[gl] != ol).sum()
    assert mism / gl.size < 0.002, mism


def test_palette_golden_within_one_lsb(oracle, tokyo):
    # samples.sh:5 — strip of 8 swatches of 40 px sorted by L
    pal = oracle.palette(tokyo, 8)
    strip = load_rgba("tokyo-palette-c8-kmeans-s40.png")
    gold = np.array([strip[20, 20 + 40 * i] for i in range(8)]).astype(int)
    assert np.abs(pal[:, :3].astype(int) - gold[:, :3]).max() <= 1
    assert hexes(pal) == ["#12110E", "#2E221E", "#602B1C", "#515346", "#AF2C1B", "#CC7550", "#869891", "#E0E2D7"]
