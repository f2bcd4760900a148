"""Multi-GPU and host-driver GPU tests.
 - the MCU-row partition of one image across the box's GPUs -- the C entry jpgb200_mgpu_encode (two
   ncclAllGathers, segments stored straight into rank 0's buffer through a peer-mapped pointer) and the
   earlier torch.distributed orchestration -- must produce the very bytes of a single-GPU encode and of the
   oracle's bit writer (needs >= 2 GPUs: skipped on the 1-GPU box, run with `gpurun --gpus 2`)
 - the host C mirror: the reference-style driver executable and the staged JpgData API."""
import io
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "jpeg-encoder-and-decoder_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _ngpu():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("w,h,q,ratio", [(512, 256, 75, 0), (1024, 1032, 90, 1), (4096, 4096, 75, 1)])
def test_partitioned_encode_equals_single_gpu(tmp_path, w, h, q, ratio):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    out = str(tmp_path / "p.npy")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", "29617", os.path.join(ROOT, "tests", "partition_worker.py"), str(w), str(h), str(q), str(ratio), out]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=280)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "identical=True" in res.stdout and "c_nccl=True" in res.stdout and "torch_dist=True" in res.stdout
    jf = np.load(out)
    from PIL import Image

    im = Image.open(io.BytesIO(jf.tobytes()))
    im.load()
    assert im.size == (w, h)
    if w * h <= 1024 * 1032:
        # the stitched stream against the ORACLE: the whole image's bytes from the oracle's bit writer fed with
        # the oracle's own coefficients differ only where FP32 and FP64 coefficients differ (<= 0.1 %), so compare
        # the decoded pictures and, exactly, the stream re-coded from this image's GPU coefficients
        r, g, b = O.synth_planes(w, h, 1005, 0, 0)
        import __graft_entry__ as ge

        e1 = ge.load_package().Encoder(0)
        try:
            zz = e1.coefficients(r, g, b, w, h, q, ratio)
        finally:
            e1.close()
        scan, _ = O.encode_scan(zz, w, h, ratio)
        want = np.concatenate([O.jfif_header(w, h, q, ratio), scan, np.array([0xFF, 0xD9], np.uint8)])
        assert jf.shape == want.shape and (jf == want).all()


def test_driver_executable_known_answer_and_file_encode(tmp_path):
    """jpeg-encoder-and-decoder_b200/jpg: no arguments = the reference driver's test_dct() walk
    (every stage on the GPU), printed in the reference's format -> must equal the golden stdout
    captured from the compiled reference; with arguments it encodes a BMP to a JPEG file."""
    exe = os.path.join(PKG, "jpg")
    got = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert got.returncode == 0, got.stderr
    def nums(text):
        out = []
        for t in text.split():
            try:
                out.append(float(t))
            except ValueError:
                pass
        return np.array(out)
    a, b = nums(got.stdout), nums(open(os.path.join(GOLDEN, "test_dct_stdout.txt")).read())
    assert a.shape == b.shape == (256,)
    assert np.abs(a - b).max() <= 0.011  # two printed decimals
    assert (a[192:] == b[192:]).all()    # the zig-zag integers: exact
    # file encode through encode_bmp_to_jpeg (fused path) and through the staged JpgData API
    for extra, name in ([], "fused.jpg"), (["--staged"], "staged.jpg"):
        outp = str(tmp_path / name)
        r = subprocess.run([exe, os.path.join(GOLDEN, "tiger.bmp"), outp, "75", "0"] + extra, capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr
    fused = np.fromfile(str(tmp_path / "fused.jpg"), np.uint8)
    staged = np.fromfile(str(tmp_path / "staged.jpg"), np.uint8)
    r_, g_, b_, w, h = O.bmp_load(os.path.join(GOLDEN, "tiger.bmp"))
    from PIL import Image

    for jf in (fused, staged):
        im = Image.open(io.BytesIO(jf.tobytes()))
        im.load()
        assert im.size == (w, h)
    # staged mode computes the coefficients in FP64 with round(): its stream is the oracle's, bit for bit
    want = O.encode_jfif(r_, g_, b_, w, h, 75, 0)
    assert staged.shape == want.shape and (staged == want).all()
    # the fused path may differ from it only through the +-1 LSB coefficient tolerance
    assert abs(int(fused.size) - int(want.size)) < 64


def test_reference_driver_source_runs_on_the_product():
    """oracle/_ref/ref_driver_b200 = the reference's own, unmodified src/jpg_driver.c linked with libjpgenc.so +
    libjpgb200.so (oracle/Makefile `driver`): its test_dct() walk runs every stage on the GPU and must print what
    the all-reference build printed (tests/golden/test_dct_stdout.txt)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_driver_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_driver_b200 not built (no reference sources in the build container)")
    got = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert got.returncode == 0, got.stderr

    def nums(text):
        out = []
        for t in text.split():
            try:
                out.append(float(t))
            except ValueError:
                pass
        return np.array(out)

    a, b = nums(got.stdout), nums(open(os.path.join(GOLDEN, "test_dct_stdout.txt")).read())
    assert a.shape == b.shape == (256,)
    assert np.abs(a - b).max() <= 0.011  # two printed decimals
    assert (a[192:] == b[192:]).all()    # the zig-zag integers: exact


def test_staged_stage_kernels_match_reference_stages(enc):
    """GPU-backed preprocess / subsample / dpcm / histogram stages vs the golden dumps."""
    d = np.load(os.path.join(GOLDEN, "tiger_q50.npz"))
    r, g, b, w, h = O.bmp_load(os.path.join(GOLDEN, "tiger.bmp"))
    enc.set_oob_bytes(d["oob"])
    try:
        Y, Cb, Cr = enc.stage_preprocess(r, g, b, w, h)
    finally:
        enc.set_oob_bytes(np.zeros(24, np.uint8))
    nb = Y.shape[0]
    blocks = np.concatenate([Y, Cb, Cr])
    dct = enc.stage_dct(blocks)
    _, lum, chr_ = O.quant_tables(50)
    qy = enc.stage_quantise(dct[:nb], lum)
    qc = enc.stage_quantise(dct[nb:], chr_)
    zz = enc.stage_zigzag(np.concatenate([qy, qc]))
    assert (zz == d["zz"]).all(), "FP64 staged pipeline must reproduce the reference's coefficients exactly"
    for c in range(3):
        dd = enc.stage_dpcm(zz[c * nb:(c + 1) * nb, 0])
        assert (dd == d["dpcm"][c]).all()
        zz[c * nb:(c + 1) * nb, 0] = dd
    fdc, fac = enc.stage_histograms(zz[:nb], O.SYM_REF)
    assert (fdc[:256] == d["freq"][0][:256]).all() and (fac[:256] == d["freq"][1][:256]).all()
    fdc, fac = enc.stage_histograms(zz[nb:], O.SYM_REF)
    assert (fdc[:256] == d["freq"][2][:256]).all() and (fac[:256] == d["freq"][3][:256]).all()
    # 4:2:2 block averaging
    half = enc.stage_subsample422(Cb, w // 8, h // 8)
    cb_img = Cb.reshape(h // 8, w // 8, 8, 8).transpose(0, 2, 1, 3).reshape(h, w)
    want = 0.5 * (cb_img[:, 0::2] + cb_img[:, 1::2])
    got = half.reshape(h // 8, w // 16, 8, 8).transpose(0, 2, 1, 3).reshape(h, w // 2)
    assert np.abs(got - want).max() < 1e-12


def test_encode_coefficients_is_byte_exact_given_equal_coefficients(enc, pkg):
    """`Given identical coefficients, the entropy-coded scan data and full JFIF byte stream are
    byte-identical` (north_star): feed the ORACLE's coefficients to the GPU entropy stage."""
    for (w, h, q, ratio, kind) in [(640, 480, 75, 0, 0), (512, 512, 97, 1, 1), (264, 40, 24, 0, 0)]:
        r, g, b = O.synth_planes(w, h, 41, 0, kind)
        zz = O.coefficients(r, g, b, w, h, q, ratio)
        want = O.encode_jfif(r, g, b, w, h, q, ratio)
        for staged in ((True, False) if pkg.binding.has_ab_forms() else (True,)):
            enc.set_staged_entropy(staged)
            got = enc.encode_coefficients(zz, w, h, q, ratio)
            assert got.shape == want.shape and (got == want).all()
        enc.set_staged_entropy(True)


def test_driver_robust_front_door(tmp_path):
    """jpg <in.bmp> <out.jpg> q ratio --robust: strict BMP parsing (data offset, 4-byte row pitch,
    bottom-up rows, B,G,R order), any size, standards-conforming arithmetic: PIL reads the BMP and
    the JPEG made from it as (nearly) the same picture -- for a width that is neither a multiple
    of 4 bytes per row nor of 8 pixels."""
    from PIL import Image

    w, h = 150, 101
    yy, xx = np.mgrid[0:h, 0:w]
    rng = np.random.default_rng(5)
    rgb = np.stack([xx * 255 // (w - 1), yy * 255 // (h - 1), 255 - (xx + yy) * 255 // (w + h - 2)], axis=-1)
    rgb = np.clip(rgb + rng.integers(-5, 6, size=rgb.shape), 0, 255).astype(np.uint8)
    bmp = str(tmp_path / "in.bmp")
    Image.fromarray(rgb, "RGB").save(bmp)  # bottom-up, rows padded to 452 bytes
    exe = os.path.join(PKG, "jpg")
    for ratio in (0, 1, 2):
        outp = str(tmp_path / ("out%d.jpg" % ratio))
        r = subprocess.run([exe, bmp, outp, "90", str(ratio), "--robust"], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr + r.stdout
        im = Image.open(outp)
        assert im.size == (w, h)
        dec = np.asarray(im.convert("RGB")).astype(np.float64)
        mse = float(((dec - rgb.astype(np.float64)) ** 2).mean())
        assert 10 * np.log10(255.0 ** 2 / mse) > 32.0
