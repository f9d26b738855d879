"""Deterministic test inputs shared by the CPU and GPU suites (numpy only)."""
from __future__ import annotations

import numpy as np


def adversarial(width_blocks: int = 16) -> np.ndarray:
    """A strip of 8-row bands, each stressing one thing:
    constants, extremes, checkerboards, ramps, impulses, exact .5 quantisation ties,
    smooth low-pass content, and noise.  float32, values 0..255 (integers)."""
    W = width_blocks * 8
    yy, xx = np.mgrid[0:8, 0:W]
    bands = [
        np.zeros((8, W)), np.full((8, W), 255.0), np.full((8, W), 128.0), np.full((8, W), 127.0),
        ((xx + yy) % 2) * 255.0,                      # pixel checkerboard
        (((xx // 8) + 0) % 2) * 255.0,                # block checkerboard
        (xx * 255.0 / (W - 1)).round(),               # horizontal ramp
        (yy * 255.0 / 7).round() + 0 * xx,            # vertical ramp
        np.where((xx % 8 == 0) & (yy == 0), 255.0, 0.0),   # DC-corner impulses
        np.where((xx % 8 == 7) & (yy == 7), 255.0, 0.0),
        (128 + 100 * np.cos(np.pi * (2 * (xx % 8) + 1) / 16)).round(),   # one horizontal basis fn
        (128 + 100 * np.cos(np.pi * (2 * yy + 1) * 3 / 16)).round() + 0 * xx,
        (128 + 60 * np.sin(xx / 9.0) + 40 * np.cos(yy / 3.0)).round(),   # smooth
    ]
    # blocks whose DC lands exactly on a rounding tie: DC = 8*(mean-128); Q[0][0]=16 ->
    # ties when 8*(mean-128) = 16*(k+0.5), i.e. mean-128 = 2k+1  (constant blocks 129,131,..)
    tie = np.zeros((8, W))
    for b in range(width_blocks):
        tie[:, b * 8:(b + 1) * 8] = 128 + (2 * (b - width_blocks // 2) + 1)
    bands.append(tie)
    rng = np.random.default_rng(1234)
    bands.append(rng.integers(0, 256, (8, W)).astype(np.float64))
    bands.append(rng.integers(120, 137, (8, W)).astype(np.float64))    # near-grey noise: many ties
    return np.clip(np.concatenate(bands, 0), 0, 255).astype(np.float32)


def float_noise(H: int, W: int, seed: int = 7) -> np.ndarray:
    """Non-integer pixels (the API takes arbitrary floats; the reference does too)."""
    rng = np.random.default_rng(seed)
    return (rng.random((H, W), dtype=np.float32) * 300.0 - 20.0).astype(np.float32)


def smooth_image(H: int, W: int) -> np.ndarray:
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    v = 128 + 70 * np.sin(xx / 37.0) * np.cos(yy / 23.0) + 30 * np.sin((xx + yy) / 11.0)
    return np.clip(v.round(), 0, 255).astype(np.float32)


def splitmix_u8(n: int, seed: int = 42, offset: int = 0) -> np.ndarray:
    """Counter-based generator used for the large synthetic images (same formula as the
    device-side generator in bench.py): v = splitmix64(seed + index) & 255."""
    idx = (np.arange(offset, offset + n, dtype=np.uint64) + np.uint64(seed)) * np.uint64(0x9E3779B97F4A7C15)
    z = idx
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    z = z ^ (z >> np.uint64(31))
    return (z & np.uint64(255)).astype(np.uint8)
