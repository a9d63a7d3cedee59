//! Host-side colour conversions through the `palette` crate, exactly where the reference does
//! them on the host too: fixed palettes on their way to the device
//! (`CentroidsBuffer::fixed_centroids`, core/src/structures.rs:523-553), centroids on their way
//! back (`CentroidsBuffer::pull_values`, :600-617) and the lightness sort key of the palettes
//! (core/src/lib.rs:276-284, 318-329).  Centroids cross the C ABI as `k x [f32; 4]`.
use palette::{IntoColor, Lab, Srgb, Srgba};
use rgb::RGBA8;

use crate::ColorSpace;

/// sRGB8 colours -> `k x [c0, c1, c2, 1.0]` in `color_space` units.
pub(crate) fn fixed_centroids(colors: &[RGBA8], color_space: &ColorSpace) -> Vec<f32> {
    let mut data = Vec::with_capacity(colors.len() * 4);
    for c in colors {
        let srgb: Srgb<f32> = Srgb::new(c.r, c.g, c.b).into_format();
        match color_space {
            ColorSpace::Lab => {
                let lab: Lab = srgb.into_color();
                data.extend_from_slice(&[lab.l, lab.a, lab.b, 1.0]);
            }
            ColorSpace::Rgb => data.extend_from_slice(&[srgb.red, srgb.green, srgb.blue, 1.0]),
        }
    }
    data
}

/// `k x [f32; 4]` centroids -> sRGB8 (alpha 255).
pub(crate) fn centroids_to_rgba8(centroids: &[f32], color_space: &ColorSpace) -> Vec<RGBA8> {
    centroids
        .chunks_exact(4)
        .map(|c| {
            let raw: Srgba<u8> = match color_space {
                ColorSpace::Lab => {
                    let s: Srgba = Lab::new(c[0], c[1], c[2]).into_color();
                    s.into_format()
                }
                ColorSpace::Rgb => Srgba::new(c[0], c[1], c[2], 1.0).into_format(),
            };
            RGBA8 {
                r: raw.red,
                g: raw.green,
                b: raw.blue,
                a: raw.alpha,
            }
        })
        .collect()
}

fn lightness(c: &RGBA8) -> f32 {
    let lab: Lab = Srgba::new(c.r, c.g, c.b, c.a)
        .into_format::<f32, f32>()
        .into_color();
    lab.l
}

/// Palettes are returned darkest first (a NaN lightness cannot come out of 8-bit input).
pub(crate) fn sort_by_lightness(colors: &mut [RGBA8]) {
    colors.sort_unstable_by(|x, y| lightness(x).partial_cmp(&lightness(y)).unwrap());
}
