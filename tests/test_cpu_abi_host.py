"""CPU tests (no GPU): the C-ABI library loads and exports every symbol include/jpgb200.h
declares, its host-only entry points agree with the oracle, the host C mirror's BMP loader
reproduces the reference loader's quirks, and the product sources never touch oracle/."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "jpeg-encoder-and-decoder_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge

    pkg = ge.load_package()
    pkg.build.build_all()
    return pkg


def header_functions():
    text = open(os.path.join(ROOT, "include", "jpgb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(jpgb200_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built):
    names = header_functions()
    assert len(names) >= 40
    lib = built.binding.lib()
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, "declared in include/jpgb200.h but not exported: %s" % missing
    out = subprocess.check_output(["nm", "-D", "--defined-only", built.binding.LIB_PATH], text=True)
    exported = set(re.findall(r"\bT (jpgb200_\w+)", out))
    assert set(names) <= exported
    assert set(built.binding.exported_names()) == set(names), "binding.py and the header disagree"


def test_no_device_means_error_not_fallback(built):
    lib = built.binding.lib()
    if lib.jpgb200_device_count() > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert lib.jpgb200_create(C.byref(h), 0) == built.binding.ERR_NO_DEVICE
    with pytest.raises(built.Jpgb200Error):
        built.Encoder(0)


def test_host_only_entry_points_match_oracle(built):
    B = built.binding
    for (w, h, q, r) in [(64, 64, 75, 0), (1920, 1080, 90, 1), (7680, 4320, 75, 0), (65520, 8, 24, 0), (16, 8, 97, 1)]:
        assert (B.jfif_header(w, h, q, r) == O.jfif_header(w, h, q, r)).all()
    for q in (24, 50, 75, 90, 97):
        rc, lum, chr_ = B.quant_tables(q)
        _, ol, oc = O.quant_tables(q)
        assert rc == 0 and (lum == ol).all() and (chr_ == oc).all()
    assert B.quant_tables(98)[0] == B.ERR_BAD_QUALITY and B.quant_tables(23)[0] == B.ERR_BAD_QUALITY
    assert B.num_blocks(7680, 4320, 0) == [518400, 518400, 518400]
    assert B.num_blocks(3840, 2160, 1) == [129600, 64800, 64800]
    for bad in [(12, 8, 0), (8, 8, 1), (8, 12, 0), (8, 8, 2)]:
        with pytest.raises(built.Jpgb200Error):
            B.num_blocks(*bad)
    lib = B.lib()
    assert b"no CPU fallback" in lib.jpgb200_strerror(B.ERR_NO_DEVICE)
    assert lib.jpgb200_kernel_name(0) == b"k1_transform"


def test_host_bmp_loader_reproduces_reference_quirks(built, tmp_path):
    """host/bitmap.c vs the oracle loader (itself pinned to the reference's planes by sha256 in
    test_cpu_oracle): reference images and a synthetic file with trailing bytes."""
    L = C.CDLL(os.path.join(PKG, "libjpgenc.so"))
    L.bmp_OpenBitmap.restype = C.c_void_p
    L.bmp_OpenBitmap.argtypes = [C.c_char_p]
    for f in ("bmp_GetRed", "bmp_GetGreen", "bmp_GetBlue"):
        getattr(L, f).restype = C.POINTER(C.c_uint8)
        getattr(L, f).argtypes = [C.c_void_p]
    for f in ("bmp_GetWidth", "bmp_GetHeight", "bmp_GetNumPixels", "bmp_GetErrorCode", "bmp_GetFileSize"):
        getattr(L, f).argtypes = [C.c_void_p]
    L.bmp_DestroyBitmap.argtypes = [C.c_void_p]
    syn = str(tmp_path / "syn.bmp")
    r, g, b = O.synth_planes(64, 32, 5, 0, 0)
    O.write_bmp(syn, r, g, b, 64, 32, trailing=2)
    for path in (os.path.join(GOLDEN, "cam.bmp"), os.path.join(GOLDEN, "tiger.bmp"), syn):
        want = O.bmp_load(path)
        bmp = L.bmp_OpenBitmap(path.encode())
        assert L.bmp_GetErrorCode(bmp) == 0
        w, h = L.bmp_GetWidth(bmp), L.bmp_GetHeight(bmp)
        assert (w, h) == (want[3], want[4]) and L.bmp_GetNumPixels(bmp) == w * h
        assert L.bmp_GetFileSize(bmp) == os.path.getsize(path)
        for i, f in enumerate(("bmp_GetRed", "bmp_GetGreen", "bmp_GetBlue")):
            got = np.ctypeslib.as_array(getattr(L, f)(bmp), shape=(w * h,))
            assert (got == want[i]).all()
        L.bmp_DestroyBitmap(bmp)
    bmp = L.bmp_OpenBitmap(b"/nonexistent/x.bmp")  # the reference crashes here (fclose(NULL))
    assert L.bmp_GetErrorCode(bmp) == 1
    L.bmp_DestroyBitmap(bmp)


def test_host_mirror_exports_reference_api(built):
    out = subprocess.check_output(["nm", "-D", "--defined-only", os.path.join(PKG, "libjpgenc.so")], text=True)
    have = set(re.findall(r"\bT (\w+)", out))
    ref_api = {"encode_bmp_to_jpeg", "encode_rgb_to_jpeg", "preprocess_jpeg", "chroma_subsample", "dct", "dct_block",
               "quantise", "quantise_lum", "quantise_chr", "zig_zag", "zig_zag_block", "dpcm", "huffman_encode",
               "calculate_freq_block_DC", "calculate_freq_block_AC", "new_block", "get_value_block", "set_value_block",
               "copy_block", "show_block", "destroy_block", "bmp_OpenBitmap", "bmp_GetRed", "bmp_GetGreen", "bmp_GetBlue",
               "bmp_GetWidth", "bmp_GetHeight", "bmp_GetNumPixels", "bmp_ShowBmpInfo", "bmp_GetFileSize",
               "bmp_GetLastError", "bmp_DestroyBitmap"}
    assert ref_api <= have, ref_api - have


def test_product_never_touches_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use oracle/"""
    for base, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".c", ".py")):
                text = open(os.path.join(base, f)).read()
                assert "oracle_lib" not in text and "libjpgb200_oracle" not in text and "orc_" not in text, f
    text = open(os.path.join(ROOT, "include", "jpgb200.h")).read()
    assert "orc_" not in text


def test_reference_driver_source_links_against_the_product(built):
    """The drop-in claim at the source level: the reference's OWN src/jpg_driver.c, unmodified and compiled
    where it lies, builds against host/headers and links with libjpgenc.so + libjpgb200.so in place of the
    reference's stage files (oracle/Makefile target `driver`).  Without a GPU the binary must stop with the
    library's "no CPU fallback" message instead of computing anything; the GPU run is in tests/test_gpu_multi.py."""
    ref_src = "/root/reference/src/jpg_driver.c"
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_driver_b200")
    if os.path.exists(ref_src):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "driver"])
    if not os.path.exists(exe):
        pytest.skip("reference sources not present and no prebuilt oracle/_ref/ref_driver_b200")
    undefined = subprocess.check_output(["nm", "-D", "--undefined-only", exe], text=True)
    for sym in ("dct_block", "quantise_lum", "zig_zag_block", "new_block", "set_value_block"):
        assert sym in undefined, "the driver must take %s from the product library" % sym
    ldd = subprocess.check_output(["ldd", exe], text=True)
    assert "libjpgenc.so" in ldd and "libjpgb200.so" in ldd and "not found" not in ldd
    import ctypes

    L = ctypes.CDLL(os.path.join(PKG, "libjpgb200.so"))
    if L.jpgb200_device_count() == 0:
        r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
        assert "no CPU fallback" in (r.stdout + r.stderr)
