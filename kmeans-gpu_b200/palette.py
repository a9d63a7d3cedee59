"""Palette parsing rules of the reference CLI (cli/src/args.rs:160-231).

The CLI itself stays unchanged in the reference; these helpers exist so non-Rust callers feed the
hot path palettes in exactly the same order (the dither result depends on palette order,
SURVEY.md R9).
"""
from __future__ import annotations

import re
from pathlib import Path

import numpy as np

_PALETTE_RE = re.compile(r"^#[0-9a-fA-F]{6}(?:,#[0-9a-fA-F]{6})*$")


def parse_colors(colors: str) -> np.ndarray:
    """cli/src/args.rs:218-231 — "#RRGGBB,#RRGGBB" in the order typed, alpha 255."""
    out = []
    for c in colors.split(","):
        out.append([int(c[1:3], 16), int(c[3:5], 16), int(c[5:7], 16), 255])
    return np.array(out, dtype=np.uint8).reshape(-1, 4)


def parse_palette(path) -> np.ndarray:
    """cli/src/args.rs:197-216 — palette image of at most 512 pixels, sorted as RGBA tuples;
    repeated colours are an error."""
    from PIL import Image as PILImage

    img = np.array(PILImage.open(path).convert("RGBA"))
    pixels = img.reshape(-1, 4)
    if pixels.shape[0] > 512:
        raise ValueError("Trying to load a palette with more than 512 colors")
    uniq = sorted(set(map(tuple, pixels.tolist())))
    if len(uniq) < pixels.shape[0]:
        raise ValueError("Trying to load a palette with recuring colors")
    return np.array(uniq, dtype=np.uint8).reshape(-1, 4)


def validate_palette(s: str) -> np.ndarray:
    """cli/src/args.rs:181-195."""
    if _PALETTE_RE.match(s):
        return parse_colors(s)
    p = Path(s)
    if len(s) > 4 and (s.endswith(".png") or s.endswith(".jpg")) and p.exists():
        return parse_palette(p)
    raise ValueError('The palette should be a path to an image file, or defined as "#RRGGBB,#RRGGBB,#RRGGBB"')
