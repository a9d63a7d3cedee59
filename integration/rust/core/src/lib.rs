//! kmeans-color-gpu on CUDA (B200, sm_100a): the public API of the crate is unchanged — same types,
//! same method signatures, same results — but everything below it (wgpu device / textures / bind
//! groups, the WGSL shaders and their preprocessor) is replaced by calls into libkmeans_gpu.so,
//! built by `build.rs` with nvcc.  `image.rs` and `octree.rs` are the reference's files, untouched;
//! `operations.rs`, `modules.rs`, `structures.rs`, `future.rs`, `utils.rs`, `shader_tests.rs` and
//! `shaders/` are gone.
use std::{fmt, ptr, str::FromStr};

use anyhow::{anyhow, ensure, Result};
pub use rgb::RGBA8;

use crate::image::{Container, Image};
use crate::octree::ColorTree;

mod colors;
mod ffi;
mod octree;

pub mod image;

/// Handle on one CUDA device.  `kmg_ctx` serialises nothing: every call takes its own stream and
/// scratch from a pool inside the context, so one processor can be shared through `Arc` by many
/// threads (examples/parallel.rs) exactly like the wgpu-based one.
pub struct ImageProcessor {
    ctx: *mut ffi::KmgCtx,
}

// SAFETY: the C library guards its per-context state (workspace pool, constant-bank slots) itself
// and keeps its error string per thread.
unsafe impl Send for ImageProcessor {}
unsafe impl Sync for ImageProcessor {}

impl Drop for ImageProcessor {
    fn drop(&mut self) {
        if !self.ctx.is_null() {
            unsafe { ffi::kmg_destroy(self.ctx) };
        }
    }
}

fn pixel_bytes(pixels: &[RGBA8]) -> *const u8 {
    pixels.as_ptr().cast()
}

fn check_image<C: Container>(image: &Image<C>) -> Result<(u32, u32)> {
    let (w, h) = image.dimensions;
    ensure!(
        image.rgba.len() as u64 == w as u64 * h as u64,
        "image is {w}x{h} but holds {} pixels",
        image.rgba.len()
    );
    Ok((w, h))
}

impl ImageProcessor {
    /// Binds CUDA device `KMG_DEVICE` (default 0).  There is no CPU or wgpu fallback: without a
    /// B200 this returns the library's error.
    /// ```rust,no_run
    /// use pollster::FutureExt;
    /// use kmeans_color_gpu::ImageProcessor;
    ///
    /// let image_processor = ImageProcessor::new().block_on();
    /// ```
    pub async fn new() -> Result<Self> {
        let abi = unsafe { ffi::kmg_abi_version() };
        ensure!(
            abi == ffi::KMG_ABI_VERSION,
            "libkmeans_gpu.so has ABI version {abi}, this crate was written against {}",
            ffi::KMG_ABI_VERSION
        );
        let device = match std::env::var("KMG_DEVICE") {
            Ok(v) => v.parse::<i32>().map_err(|e| anyhow!("KMG_DEVICE={v}: {e}"))?,
            Err(_) => 0,
        };
        let mut ctx = ptr::null_mut();
        ffi::check(unsafe { ffi::kmg_create(device, &mut ctx) })?;
        Ok(Self { ctx })
    }

    pub async fn palette<C: Container>(
        &self,
        color_count: u32,
        image: &Image<C>,
        algo: Algorithm,
    ) -> Result<Vec<RGBA8>> {
        match algo {
            Algorithm::Kmeans => kmeans_palette(self, color_count, image).await,
            Algorithm::Octree => octree_palette(self, color_count, image).await,
        }
    }

    pub async fn find<C: Container>(
        &self,
        image: &Image<C>,
        colors: &[RGBA8],
        reduce_mode: &ReduceMode,
    ) -> Result<Image<Vec<RGBA8>>> {
        let (w, h) = check_image(image)?;
        let centroids = colors::fixed_centroids(colors, &ColorSpace::Lab);
        self.remap(image, (w, h), &centroids, reduce_mode)
    }

    pub async fn reduce<C: Container>(
        &self,
        color_count: u32,
        image: &Image<C>,
        algo: &Algorithm,
        reduce_mode: &ReduceMode,
    ) -> Result<Image<Vec<RGBA8>>> {
        let (w, h) = check_image(image)?;
        match algo {
            Algorithm::Kmeans => {
                // shrink -> Lab -> farthest-point init -> Lloyd loop -> remap in one call: the image
                // goes up once and the centroids never leave the device
                let mut out = vec![RGBA8::default(); image.rgba.len()];
                ffi::check(unsafe {
                    ffi::kmg_reduce(
                        self.ctx,
                        pixel_bytes(&image.rgba),
                        w,
                        h,
                        color_count,
                        ColorSpace::Lab as i32,
                        *reduce_mode as i32,
                        ptr::null(), // reference constants
                        out.as_mut_ptr().cast(),
                        ptr::null_mut(),
                        ptr::null_mut(),
                    )
                })?;
                Ok(Image::new((w, h), out))
            }
            Algorithm::Octree => {
                let palette = octree_palette(self, color_count, image).await?;
                let centroids = colors::fixed_centroids(&palette, &ColorSpace::Lab);
                self.remap(image, (w, h), &centroids, reduce_mode)
            }
        }
    }

    /// `find_colors` / `dither_colors` / `meld_colors` + `pull_image` of the reference in one call.
    fn remap<C: Container>(
        &self,
        image: &Image<C>,
        (w, h): (u32, u32),
        centroids: &[f32],
        reduce_mode: &ReduceMode,
    ) -> Result<Image<Vec<RGBA8>>> {
        let mut out = vec![RGBA8::default(); image.rgba.len()];
        ffi::check(unsafe {
            ffi::kmg_remap(
                self.ctx,
                pixel_bytes(&image.rgba),
                w,
                h,
                centroids.as_ptr(),
                (centroids.len() / 4) as u32,
                ColorSpace::Lab as i32,
                *reduce_mode as i32,
                out.as_mut_ptr().cast(),
            )
        })?;
        Ok(Image::new((w, h), out))
    }
}

/// Discriminants are `kmg_color_space`.
#[derive(Clone, Copy)]
#[repr(i32)]
pub enum ColorSpace {
    Lab = 0,
    Rgb = 1,
}

impl ColorSpace {
    pub fn from(str: &str) -> Option<ColorSpace> {
        str.parse().ok()
    }

    pub fn name(&self) -> &'static str {
        match self {
            ColorSpace::Lab => "lab",
            ColorSpace::Rgb => "rgb",
        }
    }

    /// Movement (CIE94) below which a centroid counts as converged; `kmg_opts::convergence < 0`
    /// selects the same values inside the library.
    pub fn convergence(&self) -> f32 {
        match self {
            ColorSpace::Lab => 1.0,
            ColorSpace::Rgb => 0.01,
        }
    }
}

impl FromStr for ColorSpace {
    type Err = anyhow::Error;

    fn from_str(s: &str) -> Result<Self, Self::Err> {
        [ColorSpace::Lab, ColorSpace::Rgb]
            .into_iter()
            .find(|c| c.name() == s)
            .ok_or_else(|| anyhow!("Unsupported color space {s}"))
    }
}

impl fmt::Display for ColorSpace {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        f.write_str(self.name())
    }
}

#[derive(Clone, Copy)]
pub enum Algorithm {
    Kmeans,
    Octree,
}

impl fmt::Display for Algorithm {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        f.write_str(match self {
            Algorithm::Kmeans => "kmeans",
            Algorithm::Octree => "octree",
        })
    }
}

/// Discriminants are `kmg_reduce_mode`.
#[derive(Clone, Copy)]
#[repr(i32)]
pub enum ReduceMode {
    Replace = 0,
    Dither = 1,
    Meld = 2,
}

impl fmt::Display for ReduceMode {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        f.write_str(match self {
            ReduceMode::Replace => "replace",
            ReduceMode::Dither => "dither",
            ReduceMode::Meld => "meld",
        })
    }
}

/// `extract_palette_kmeans` + `pull_values` + the sort by lightness.
async fn kmeans_palette<C: Container>(
    image_processor: &ImageProcessor,
    color_count: u32,
    image: &Image<C>,
) -> Result<Vec<RGBA8>> {
    let (w, h) = check_image(image)?;
    let mut centroids = vec![0f32; color_count as usize * 4];
    ffi::check(unsafe {
        ffi::kmg_kmeans_palette(
            image_processor.ctx,
            pixel_bytes(&image.rgba),
            w,
            h,
            color_count,
            ColorSpace::Lab as i32,
            ptr::null(),
            centroids.as_mut_ptr(),
            ptr::null_mut(),
        )
    })?;
    let mut colors = colors::centroids_to_rgba8(&centroids, &ColorSpace::Lab);
    colors::sort_by_lightness(&mut colors);
    Ok(colors)
}

/// The octree quantiser stays on the CPU (octree.rs, untouched); only the bilinear shrink to at
/// most 128 px in front of it ran on the GPU in the reference, and still does (`kmg_resize`).
async fn octree_palette<C: Container>(
    image_processor: &ImageProcessor,
    color_count: u32,
    image: &Image<C>,
) -> Result<Vec<RGBA8>> {
    const MAX_SIZE: u32 = 128;

    let (w, h) = check_image(image)?;
    let shrunk: Option<Vec<RGBA8>> = if w > MAX_SIZE || h > MAX_SIZE {
        let (mut sw, mut sh) = (0u32, 0u32);
        unsafe { ffi::kmg_resized_dims(w, h, MAX_SIZE, &mut sw, &mut sh) };
        let mut small = vec![RGBA8::default(); sw as usize * sh as usize];
        ffi::check(unsafe {
            ffi::kmg_resize(
                image_processor.ctx,
                pixel_bytes(&image.rgba),
                w,
                h,
                MAX_SIZE,
                small.as_mut_ptr().cast(),
            )
        })?;
        Some(small)
    } else {
        None
    };
    let pixels: &[RGBA8] = shrunk.as_deref().unwrap_or(&image.rgba[..]);

    let mut tree = ColorTree::new();
    for pixel in pixels {
        tree.add_color(pixel);
    }
    let mut colors = tree.reduce(color_count as usize);
    colors::sort_by_lightness(&mut colors);
    Ok(colors)
}

/// Batch callers (examples/gif.rs reduces every frame of an animation with the same settings):
/// `frames` holds `n` frames of `dimensions` back to back; one upload / kernel / read-back pipeline
/// runs over all of them.  Not part of the reference's API.
pub fn reduce_frames(
    image_processor: &ImageProcessor,
    color_count: u32,
    dimensions: (u32, u32),
    frames: &[RGBA8],
    reduce_mode: &ReduceMode,
) -> Result<Vec<RGBA8>> {
    let per_frame = dimensions.0 as usize * dimensions.1 as usize;
    ensure!(per_frame > 0 && frames.len() % per_frame == 0, "frames do not fill whole images");
    let mut out = vec![RGBA8::default(); frames.len()];
    ffi::check(unsafe {
        ffi::kmg_reduce_batch(
            image_processor.ctx,
            pixel_bytes(frames),
            (frames.len() / per_frame) as u32,
            dimensions.0,
            dimensions.1,
            color_count,
            ColorSpace::Lab as i32,
            *reduce_mode as i32,
            ptr::null(),
            out.as_mut_ptr().cast(),
            ptr::null_mut(),
            ptr::null_mut(),
        )
    })?;
    Ok(out)
}
