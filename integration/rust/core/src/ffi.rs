//! Raw bindings of `include/kmeans_gpu.h` (C ABI of libkmeans_gpu.so) — one declaration per entry
//! point this crate calls.  Field order and integer widths mirror the header; `KMG_ABI_VERSION` is
//! checked once in `ImageProcessor::new`.
#![allow(dead_code)]

use std::ffi::{c_void, CStr};
use std::os::raw::{c_char, c_int};

use anyhow::{anyhow, Result};

pub const KMG_ABI_VERSION: c_int = 1;

/// `kmg_color_space`
pub const KMG_LAB: c_int = 0;
pub const KMG_RGB: c_int = 1;
/// `kmg_reduce_mode`
pub const KMG_REPLACE: c_int = 0;
pub const KMG_DITHER: c_int = 1;
pub const KMG_MELD: c_int = 2;

/// Opaque `kmg_ctx`.
#[repr(C)]
pub struct KmgCtx {
    _opaque: [u8; 0],
}

/// `kmg_opts` — the reference's hard-coded constants made explicit (max_dim 256, max_iter 128,
/// check_every 8, ...).  Always start from `kmg_default_opts`.
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct KmgOpts {
    pub struct_size: u32,
    pub max_dim: u32,
    pub max_iter: u32,
    pub check_every: u32,
    pub convergence: f32,
    pub seed_x_frac: f32,
    pub seed_y_frac: f32,
    pub seed_x: i32,
    pub seed_y: i32,
    pub flags: u32,
}

extern "C" {
    pub fn kmg_create(device: c_int, out: *mut *mut KmgCtx) -> c_int;
    pub fn kmg_destroy(ctx: *mut KmgCtx);
    pub fn kmg_last_error() -> *const c_char;
    pub fn kmg_abi_version() -> c_int;
    pub fn kmg_default_opts(opts: *mut KmgOpts);

    pub fn kmg_kmeans_palette(
        ctx: *mut KmgCtx,
        rgba: *const u8,
        w: u32,
        h: u32,
        k: u32,
        color_space: c_int,
        opts: *const KmgOpts,
        centroids_out: *mut f32,
        passes_out: *mut u32,
    ) -> c_int;
    pub fn kmg_remap(
        ctx: *mut KmgCtx,
        rgba: *const u8,
        w: u32,
        h: u32,
        centroids: *const f32,
        k: u32,
        color_space: c_int,
        mode: c_int,
        out_rgba: *mut u8,
    ) -> c_int;
    pub fn kmg_reduce(
        ctx: *mut KmgCtx,
        rgba: *const u8,
        w: u32,
        h: u32,
        k: u32,
        color_space: c_int,
        mode: c_int,
        opts: *const KmgOpts,
        out_rgba: *mut u8,
        centroids_out: *mut f32,
        passes_out: *mut u32,
    ) -> c_int;
    pub fn kmg_resized_dims(w: u32, h: u32, max_size: u32, out_w: *mut u32, out_h: *mut u32);
    pub fn kmg_resize(
        ctx: *mut KmgCtx,
        rgba: *const u8,
        w: u32,
        h: u32,
        max_size: u32,
        out: *mut u8,
    ) -> c_int;
    pub fn kmg_reduce_batch(
        ctx: *mut KmgCtx,
        rgba: *const u8,
        n_frames: u32,
        w: u32,
        h: u32,
        k: u32,
        color_space: c_int,
        mode: c_int,
        opts: *const KmgOpts,
        out_rgba: *mut u8,
        centroids_out: *mut f32,
        passes_out: *mut u32,
    ) -> c_int;

    pub fn kmg_alloc_pinned(bytes: usize) -> *mut c_void;
    pub fn kmg_free_pinned(p: *mut c_void);

    // host-side colour helpers; this crate keeps using the `palette` crate for them (colors.rs),
    // they are bound for callers that want the library's own restatement
    pub fn kmg_fixed_centroids(colors_rgba8: *const u8, count: u32, color_space: c_int, centroids_out: *mut f32);
    pub fn kmg_centroids_to_rgba8(centroids: *const f32, count: u32, color_space: c_int, colors_out: *mut u8);
    pub fn kmg_sort_palette_by_lightness(colors_rgba8: *mut u8, count: u32);
}

/// `anyhow::Result` from a `kmg_status`: the message is the calling thread's `kmg_last_error()`.
pub fn check(code: c_int) -> Result<()> {
    if code == 0 {
        return Ok(());
    }
    let msg = unsafe {
        let p = kmg_last_error();
        if p.is_null() {
            String::new()
        } else {
            CStr::from_ptr(p).to_string_lossy().into_owned()
        }
    };
    let kind = match code {
        1 => "bad argument",
        2 => "CUDA",
        3 => "out of memory",
        4 => "NCCL",
        5 => "unsupported",
        _ => "unknown",
    };
    Err(anyhow!("kmeans_gpu ({kind}, status {code}): {msg}"))
}
