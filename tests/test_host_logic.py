"""CPU-side checks: the C-ABI library builds, loads and exports every symbol the header declares;
host helpers (palette-crate conversions, palette parsing, shrink rule, sharding) match the oracle.
No compute entry point is exercised here (that needs a GPU: tests/test_gpu_parity.py)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


def header_functions():
    text = (ROOT / "include" / "kmeans_gpu.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kmg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(native_lib):
    names = header_functions()
    assert len(names) >= 35
    for n in names:
        assert hasattr(native_lib, n), f"{n} declared in include/kmeans_gpu.h but not exported"
    assert native_lib.kmg_abi_version() == 1


def test_binding_covers_header():
    import kmeans_gpu_b200  # noqa: F401
    from kmeans_gpu_b200 import _native

    assert set(header_functions()) == set(_native.SIGNATURES)


def test_default_opts_are_reference_constants(native_lib):
    from kmeans_gpu_b200._native import KmgOpts

    o = KmgOpts()
    native_lib.kmg_default_opts(C.byref(o))
    assert (o.max_dim, o.max_iter, o.check_every) == (256, 128, 8)  # structures.rs:23, modules.rs:765-766
    assert o.convergence < 0 and (o.seed_x, o.seed_y) == (-1, -1)
    assert (o.seed_x_frac, o.seed_y_frac) == (0.5625, 0.93359375)
    assert o.struct_size == C.sizeof(KmgOpts)


def test_no_gpu_fails_loudly(native_lib):
    """Without a device kmg_create must fail with a CUDA error (no CPU fallback exists)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import kmeans_gpu_b200 as K

    with pytest.raises(K.KmgError) as e:
        K.ImageProcessor(0)
    assert e.value.code == 2 and "no CPU fallback" in e.value.message


def test_bad_args_rejected_without_gpu(native_lib):
    assert native_lib.kmg_create(0, None) == 1
    assert b"NULL" in native_lib.kmg_last_error()


@pytest.mark.parametrize("w,h", [(768, 513), (513, 768), (300, 300), (8192, 8192), (1920, 1080), (257, 1), (1, 1000), (4000, 3)])
def test_resized_dims_match_oracle(native_lib, oracle, w, h):
    import kmeans_gpu_b200 as K

    for m in (256, 128):
        assert K.resized_dims(w, h, m) == oracle.resized_dims(w, h, m)
    # structures.rs:79-89: strict `width > height`
    assert K.resized_dims(768, 513) == (256, 171)
    assert K.resized_dims(300, 300) == (256, 256)


def test_fixed_centroids_match_oracle(native_lib, oracle):
    import kmeans_gpu_b200 as K

    rng = np.random.default_rng(7)
    cols = rng.integers(0, 256, (4096, 4), dtype=np.uint8)
    cols[:, 3] = 255
    cols[:3] = [[0, 0, 0, 255], [255, 255, 255, 255], [10, 10, 10, 255]]
    ours = K.fixed_centroids(cols, K.ColorSpace.Lab)
    assert np.array_equal(ours.view(np.uint32), oracle.pal_srgb8_to_lab(cols).view(np.uint32))
    rgb = K.fixed_centroids(cols, K.ColorSpace.Rgb)
    assert np.array_equal(rgb[:, :3], cols[:, :3].astype(np.float32) / np.float32(255.0))
    assert (rgb[:, 3] == 1.0).all()


def test_centroids_to_rgba8_and_sort_match_oracle(native_lib, oracle):
    import kmeans_gpu_b200 as K

    rng = np.random.default_rng(11)
    cols = rng.integers(0, 256, (2048, 4), dtype=np.uint8)
    cols[:, 3] = 255
    lab = oracle.pal_srgb8_to_lab(cols)
    assert np.array_equal(K.centroids_to_rgba8(lab), cols)  # round trip of 8-bit colours
    lab2 = lab + rng.normal(0, 0.3, lab.shape).astype(np.float32)
    assert np.array_equal(K.centroids_to_rgba8(lab2), oracle.pal_lab_to_srgb8(lab2))
    s = K.sort_palette_by_lightness(cols[:64])
    key = oracle.pal_srgb8_to_lab(cols[:64])[:, 0]
    assert np.array_equal(s, cols[:64][np.argsort(key, kind="stable")])


def test_palette_parsing_rules():
    """cli/src/args.rs:238-293 (the reference's own CLI tests)."""
    import kmeans_gpu_b200 as K

    c = K.parse_colors("#ffffff,#000000")
    assert c.tolist() == [[255, 255, 255, 255], [0, 0, 0, 255]]
    pal = K.parse_palette(GOLDEN / "resurrect_64.png")
    assert pal.shape == (64, 4)
    assert pal.tolist() == sorted(pal.tolist())
    assert K.validate_palette("#010203").tolist() == [[1, 2, 3, 255]]
    with pytest.raises(ValueError):
        K.validate_palette("#12345")
    with pytest.raises(ValueError):
        K.validate_palette("not-a-file.png")
    with pytest.raises(ValueError):  # more than 512 pixels
        K.parse_palette(GOLDEN / "tokyo.png")


def test_enums_mirror_reference():
    import kmeans_gpu_b200 as K

    assert K.ColorSpace.from_str("lab") is K.ColorSpace.Lab and str(K.ColorSpace.Rgb) == "rgb"
    assert K.ColorSpace.Lab.convergence() == 1.0 and K.ColorSpace.Rgb.convergence() == 0.01  # lib.rs:189-194
    with pytest.raises(ValueError):
        K.ColorSpace.from_str("xyz")
    assert str(K.Algorithm.Kmeans) == "kmeans" and str(K.Algorithm.Octree) == "octree"
    assert [str(m) for m in K.ReduceMode] == ["replace", "dither", "meld"]
    img = K.Image.new((2, 1), np.array([[1, 2, 3, 4, 5, 6, 7, 8]], np.uint8))
    assert img.dimensions == (2, 1) and img.get_pixel(1, 0) == (5, 6, 7, 8)
    assert img.into_raw_pixels().shape == (2, 4)
    with pytest.raises(ValueError):
        K.Image.new((3, 3), np.zeros(8, np.uint8))


def test_shard_rules():
    import kmeans_gpu_b200 as K

    assert K.row_shards(8192, 8) == [(i * 1024, (i + 1) * 1024) for i in range(8)]
    s = K.row_shards(513, 4)
    assert s[0][0] == 0 and s[-1][1] == 513 and all(a[1] == b[0] for a, b in zip(s, s[1:]))
    f = K.frame_shards(4096, 8)
    assert [b - a for a, b in f] == [512] * 8
    assert K.frame_shards(3, 4) == [(0, 0), (0, 1), (1, 2), (2, 3)]


def test_synth_generator_is_stable(oracle):
    """Known answers of the synthetic generator so CPU and GPU inputs agree."""
    u = oracle.synth(4, seed=1)
    b = oracle.synth(4, seed=2, blobs=512, first_pixel=5)
    assert u[:, 3].tolist() == [255] * 4 and b[:, 3].tolist() == [255] * 4
    # split generation == whole generation
    whole = oracle.synth(1000, seed=3, blobs=32, frame=7)
    parts = np.concatenate([oracle.synth(400, seed=3, blobs=32, frame=7), oracle.synth(600, first_pixel=400, seed=3, blobs=32, frame=7)])
    assert np.array_equal(whole, parts)


def test_header_is_plain_c_and_links(native_lib, tmp_path):
    """The boundary is a C ABI: include/kmeans_gpu.h must compile as C99 (no C++ in the signatures)
    and a plain C caller must link against the library (host-only entry points run without a GPU)."""
    import subprocess

    src = tmp_path / "caller.c"
    src.write_text(
        '#include "kmeans_gpu.h"\n'
        "#include <stdio.h>\n"
        "int main(void) {\n"
        "  kmg_opts o; kmg_default_opts(&o);\n"
        "  uint32_t w = 0, h = 0; kmg_resized_dims(768, 513, o.max_dim, &w, &h);\n"
        "  uint8_t px[8] = {10, 10, 10, 255, 250, 250, 250, 255}, pal[8]; uint32_t n = 0;\n"
        "  int rc = kmg_octree_palette(px, 2, 2, pal, &n);\n"
        "  float lab[8]; kmg_fixed_centroids(pal, n, KMG_LAB, lab);\n"
        '  printf("%u %u %u %d %d %.3f\\n", w, h, n, rc, kmg_abi_version(), lab[0]);\n'
        "  return 0;\n"
        "}\n")
    lib_dir = ROOT / "kmeans-gpu_b200" / "lib"
    exe = tmp_path / "caller"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", f"-I{ROOT / 'include'}", str(src), "-o",
                           str(exe), f"-L{lib_dir}", "-lkmeans_gpu", f"-Wl,-rpath,{lib_dir}"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert out[:5] == ["256", "171", "2", "0", "1"]
    assert abs(float(out[5]) - 2.742) < 0.01  # L of #0a0a0a


def _c_prototypes():
    """name -> number of parameters, from include/kmeans_gpu.h."""
    text = (ROOT / "include" / "kmeans_gpu.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for name, args in re.findall(r"\b(kmg_[a-z0-9_]+)\s*\(([^;{]*)\)\s*;", text):
        args = args.strip()
        protos[name] = 0 if args in ("", "void") else len(args.split(","))
    return protos


def test_rust_shim_matches_header():
    """integration/rust cannot be compiled here (no cargo/rustc): check at least that
    ffi.rs binds only functions the header declares, with the same number of parameters, that the
    kmg_opts mirror has the header's fields in order, and that lib.rs keeps the reference's public
    items (core/src/lib.rs:24-253) and calls nothing ffi.rs does not declare."""
    rust = ROOT / "integration" / "rust" / "core"
    ffi = (rust / "src" / "ffi.rs").read_text()
    lib = (rust / "src" / "lib.rs").read_text()
    protos = _c_prototypes()
    block = ffi[ffi.index('extern "C" {'):]
    block = block[:block.index("\n}\n")]
    block = re.sub(r"//.*", "", block)
    bound = {}
    for name, args in re.findall(r"pub fn (kmg_[a-z0-9_]+)\s*\(([^)]*)\)", block, flags=re.S):
        args = [a for a in (x.strip() for x in args.split(",")) if a]
        bound[name] = len(args)
    assert len(bound) >= 15
    for name, n in bound.items():
        assert name in protos, f"ffi.rs binds {name}, which include/kmeans_gpu.h does not declare"
        assert protos[name] == n, f"{name}: {n} parameters in ffi.rs, {protos[name]} in the header"
    for need in ("kmg_create", "kmg_destroy", "kmg_kmeans_palette", "kmg_remap", "kmg_reduce", "kmg_resize", "kmg_last_error"):
        assert need in bound
    # kmg_opts field order
    header = (ROOT / "include" / "kmeans_gpu.h").read_text()
    hdr_fields = re.findall(r"^\s+(?:uint32_t|int32_t|float)\s+(\w+);", header[header.index("typedef struct kmg_opts"):header.index("} kmg_opts;")], flags=re.M)
    rs_fields = re.findall(r"pub (\w+): [a-z0-9]+,", ffi[ffi.index("pub struct KmgOpts"):ffi.index("extern")])
    assert hdr_fields == rs_fields and len(rs_fields) == 10
    # every ffi:: call in lib.rs is declared in ffi.rs
    for name in set(re.findall(r"ffi::(kmg_[a-z0-9_]+)", lib)):
        assert name in bound, name
    # the public surface of the reference crate is intact
    for item in ("pub struct ImageProcessor", "pub async fn new() -> Result<Self>", "pub async fn palette<C: Container>(",
                 "pub async fn find<C: Container>(", "pub async fn reduce<C: Container>(", "pub enum ColorSpace", "pub enum Algorithm",
                 "pub enum ReduceMode", "pub use rgb::RGBA8;", "pub mod image;", "impl FromStr for ColorSpace",
                 "pub fn convergence(&self) -> f32", "async fn kmeans_palette<C: Container>(", "async fn octree_palette<C: Container>("):
        assert item in lib, item
    assert "Replace = 0" in lib and "Dither = 1" in lib and "Meld = 2" in lib and "Lab = 0" in lib and "Rgb = 1" in lib
    build = (rust / "build.rs").read_text()
    assert "arch=compute_100a,code=sm_100a" in build and "cargo:rustc-link-lib=dylib=kmeans_gpu" in build
    csrc = {p.name for p in (ROOT / "kmeans-gpu_b200" / "csrc").glob("*.cu*")} | {"kmg_host.cpp"}
    for f in re.findall(r'"(kmg_[a-z_]+\.(?:cuh|cu|cpp))"', build):
        assert f in csrc, f"build.rs names {f}, which is not in kmeans-gpu_b200/csrc"
    assert {p.name for p in (ROOT / "kmeans-gpu_b200" / "csrc").glob("*.cuh")} <= set(re.findall(r'"(kmg_[a-z_]+\.cuh)"', build))


def test_all_16m_colours_round_trip_through_lab(native_lib):
    """R8 pin, widened: every one of the 2^24 sRGB8 colours must survive sRGB8 -> Lab (kmg_fixed_centroids,
    the `palette` crate's conversion behind CentroidsBuffer::fixed_centroids, structures.rs:523-553) -> sRGB8
    (kmg_centroids_to_rgba8, pull_values, :600-617) unchanged — the property the reference relies on when a
    fixed palette comes back out of `find`, and the tightest check available for a crate whose source is
    not in the tree (the three bit-exact goldens go through the same code)."""
    import kmeans_gpu_b200 as K

    step = 1 << 20
    for lo in range(0, 1 << 24, step):
        v = np.arange(lo, lo + step, dtype=np.uint32)
        cols = np.stack([v & 255, (v >> 8) & 255, v >> 16, np.full_like(v, 255)], axis=1).astype(np.uint8)
        lab = K.fixed_centroids(cols, K.ColorSpace.Lab)
        back = K.centroids_to_rgba8(lab)
        assert np.array_equal(back, cols), lo
