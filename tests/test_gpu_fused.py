"""GPU parity tests of K1E (csrc/k1_fused.cuh): the transform and the chunk encode in one
persistent kernel (jpgb200_set_fused_transform).  What the reference chains as
dct -> quantise -> zig_zag -> dpcm -> huffman_encode (src/jpg_encode.c:36-50) must come out

  * byte-identical to the separate kernels (K1 -> KE) and to the oracle's bit writer fed with the
    GPU's own coefficients (bit-exact bar for everything downstream of the coefficients),
  * with the same coefficients as K1 (the fused kernel's tap), within +-1 LSB on <= 0.1 % of the
    oracle's FP64 coefficients (BASELINE.json north_star).

Cases cover ragged tiles (width not a multiple of 256), a single tile, more tiles than resident
CTAs (the DC hand-over between CTAs wraps around the grid), 4:2:2, the lowest and highest quality,
noise at q97 (entries longer than their private region -> the pack-into-the-slot path), widths the
fused kernel does not take (w % 16 == 8: separate kernels) and the standards-conforming mode.
"""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu

CASES = [
    # w, h, quality, ratio, kind, seed
    (256, 8, 75, 0, 0, 101),       # exactly one tile
    (16, 8, 75, 1, 0, 102),        # a single 4:2:2 MCU
    (64, 8, 75, 0, 0, 103),        # a quarter tile
    (272, 24, 90, 1, 1, 104),      # 34 blocks per row, 4:2:2: ragged second tile
    (528, 40, 50, 0, 0, 105),      # 66 blocks per row: two full tiles and a two-block one
    (520, 72, 24, 0, 0, 106),      # w % 16 == 8: the fused kernel declines, separate kernels
    (1024, 1032, 97, 1, 1, 107),   # noise, highest quality: private regions overflow
    (1024, 512, 97, 0, 1, 108),
    (640, 480, 75, 0, 0, 109),
    (640, 480, 75, 1, 0, 110),
    (4096, 2176, 75, 0, 0, 111),   # 4352 tiles: every CTA walks ~15 tiles, the DC hand-over wraps the grid
    (4096, 1088, 90, 1, 0, 112),
    (1920, 1080, 75, 0, 0, 1001),  # BASELINE.json configs[1]
]


def oracle_stream(zz, w, h, q, ratio):
    scan, _ = O.encode_scan(zz, w, h, ratio)
    return np.concatenate([O.jfif_header(w, h, q, ratio), scan, np.array([0xFF, 0xD9], np.uint8)])


@pytest.mark.parametrize("w,h,q,ratio,kind,seed", CASES)
def test_fused_equals_separate_kernels_and_oracle(enc, w, h, q, ratio, kind, seed):
    r, g, b = O.synth_planes(w, h, seed, 0, kind)
    oob = (np.arange(24, dtype=np.uint8) * 11 + 5).astype(np.uint8)
    enc.set_oob_bytes(oob)
    try:
        enc.set_fused_transform(True)
        enc.set_fused_tap(True)
        fused = enc.encode_planar(r, g, b, w, h, q, ratio)
        zz_f = enc.last_coefficients(w, h, ratio)
        enc.set_fused_tap(False)
        fused_notap = enc.encode_planar(r, g, b, w, h, q, ratio)
        enc.set_fused_transform(False)
        sep = enc.encode_planar(r, g, b, w, h, q, ratio)
        zz = enc.coefficients(r, g, b, w, h, q, ratio)
    finally:
        enc.set_fused_transform(False)
        enc.set_fused_tap(False)
        enc.set_oob_bytes(np.zeros(24, np.uint8))
    assert (zz_f == zz).all(), "the fused kernel's coefficients differ from K1's"
    assert fused.shape == sep.shape and (fused == sep).all()
    assert fused_notap.shape == sep.shape and (fused_notap == sep).all()
    want = oracle_stream(zz, w, h, q, ratio)
    assert fused.shape == want.shape and (fused == want).all()
    if w * h <= 1920 * 1080:
        ref = O.coefficients(r, g, b, w, h, q, ratio, oob=oob)
        d = np.abs(zz.astype(np.int32) - ref.astype(np.int32))
        assert d.max() <= 1 and float((d != 0).mean()) <= 1e-3


@pytest.mark.parametrize("w,h,q,ratio,seed", [(640, 480, 75, 0, 201), (1280, 720, 90, 1, 202)])
def test_fused_standard_mode(enc, w, h, q, ratio, seed):
    """the standards-conforming front door (textbook Cb, no shifted last block column) through K1E"""
    r, g, b = O.synth_planes(w, h, seed, 0, 0)
    try:
        enc.set_standard_mode(True)
        enc.set_fused_transform(True)
        fused = enc.encode_planar(r, g, b, w, h, q, ratio)
        enc.set_fused_transform(False)
        sep = enc.encode_planar(r, g, b, w, h, q, ratio)
    finally:
        enc.set_fused_transform(False)
        enc.set_standard_mode(False)
    assert fused.shape == sep.shape and (fused == sep).all()


def test_fused_repeated_launches_and_geometry_changes(enc):
    """The DC hand-over words carry a launch epoch: alternating image sizes and repeated encodes on
    one context never see a stale word."""
    shapes = [(512, 64, 0), (256, 256, 1), (1024, 128, 0), (512, 64, 0), (256, 256, 1)]
    want = {}
    try:
        for i, (w, h, ratio) in enumerate(shapes * 3):
            r, g, b = O.synth_planes(w, h, 300 + (i % len(shapes)), 0, (i % len(shapes)) & 1)
            enc.set_fused_transform(True)
            got = enc.encode_planar(r, g, b, w, h, 75, ratio)
            key = (i % len(shapes))
            if key not in want:
                enc.set_fused_transform(False)
                want[key] = enc.encode_planar(r, g, b, w, h, 75, ratio)
            assert got.shape == want[key].shape and (got == want[key]).all(), i
    finally:
        enc.set_fused_transform(False)


@pytest.mark.parametrize("w,h,ratio,nf", [(512, 256, 0, 7), (768, 64, 1, 5), (1920, 1088, 0, 4)])
def test_fused_batch_device(enc, w, h, ratio, nf):
    """jpgb200_encode_batch_device through K1E with and without the tail overlap == per-frame encodes
    through the separate kernels."""
    q, npx = 75, w * h
    d_pl = enc.malloc(nf * 3 * npx)
    cap = 3 * npx + 4096
    d_out = enc.malloc(nf * cap)
    d_len = enc.malloc(8 * nf)
    try:
        for f in range(nf):
            enc.synth_device(d_pl + f * 3 * npx, w, h, 777, f, f & 1)
        enc.sync()
        res = {}
        enc.set_fused_transform(True)
        for mode in ("on", "off", "on2"):
            enc.set_batch_overlap(mode != "off")
            enc.h2d(d_out, np.zeros(nf * cap, np.uint8))
            enc.encode_batch_device(d_pl, nf, w, h, q, ratio, d_out, cap, d_len)
            enc.sync()
            res[mode] = (enc.d2h(d_len, 8 * nf).view(np.uint64).copy(), enc.d2h(d_out, nf * cap).reshape(nf, cap).copy())
        enc.set_batch_overlap(True)
        planes = enc.d2h(d_pl, nf * 3 * npx).reshape(nf, 3, -1)
        enc.set_fused_transform(False)
        for f in range(nf):
            one = enc.encode_planar(planes[f, 0], planes[f, 1], planes[f, 2], w, h, q, ratio)
            for mode in res:
                lens, outs = res[mode]
                assert int(lens[f]) == one.size, (mode, f)
                assert (outs[f, :one.size] == one).all(), (mode, f)
    finally:
        enc.set_fused_transform(False)
        enc.set_batch_overlap(True)
        enc.free(d_pl)
        enc.free(d_out)
        enc.free(d_len)
