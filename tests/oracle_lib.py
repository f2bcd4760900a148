"""ctypes binding of oracle/libjpgb200_oracle.so -- TEST INFRASTRUCTURE ONLY.

The oracle is the CPU restatement of the reference's encode path
(oracle/jpgb200_oracle.c).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "libjpgb200_oracle.so")
REF_HARNESS = os.path.join(ORACLE_DIR, "_ref", "ref_harness")
REF_HARNESS_O2 = os.path.join(ORACLE_DIR, "_ref", "ref_harness_O2")

DC_REF, DC_SPEC = 0, 1
SYM_REF, SYM_SPEC = 0, 1

_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "liboracle"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(
            os.path.join(ORACLE_DIR, "jpgb200_oracle.c")
        ):
            build()
        L = C.CDLL(LIB_PATH)
        u8p = C.POINTER(C.c_uint8)
        L.orc_quant_tables.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_block_counts.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_long)]
        L.orc_coefficients.argtypes = [u8p, u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_void_p, C.c_void_p]
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_set_standard.argtypes = [C.c_int]
        L.orc_dct_block.argtypes = [C.c_void_p]
        L.orc_quantise_block.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_zigzag_block.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_dc_sequence.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_void_p]
        L.orc_histograms.argtypes = [C.c_void_p, C.POINTER(C.c_long), C.c_int, C.c_int, C.c_void_p]
        L.orc_huffman_table.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.orc_encode_scan.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)]
        L.orc_encode_scan.restype = C.c_size_t
        L.orc_encode_scan_raw.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)]
        L.orc_encode_scan_raw.restype = C.c_size_t
        L.orc_block_bits.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_jfif_header.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        L.orc_jfif_header.restype = C.c_size_t
        L.orc_specs_standard.argtypes = [C.c_void_p]
        L.orc_specs_optimal.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_encode_scan_specs.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t,
                                            C.POINTER(C.c_uint64)]
        L.orc_encode_scan_specs.restype = C.c_size_t
        L.orc_jfif_header_specs.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_jfif_header_specs.restype = C.c_size_t
        L.orc_encode_jfif.argtypes = [u8p, u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_void_p, C.c_size_t]
        L.orc_encode_jfif.restype = C.c_size_t
        L.orc_bmp_load.argtypes = [C.c_char_p, C.POINTER(u8p), C.POINTER(u8p), C.POINTER(u8p), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_synth_plane.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_uint32, C.c_int, C.c_int]
        L.orc_time_encode.argtypes = [u8p, u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_time_encode.restype = C.c_double
        _lib = L
    return _lib


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _planes(r, g, b):
    return [np.ascontiguousarray(p, dtype=np.uint8).reshape(-1) for p in (r, g, b)]


def quant_tables(quality):
    lum = (C.c_int * 64)()
    chr_ = (C.c_int * 64)()
    rc = lib().orc_quant_tables(quality, lum, chr_)
    return rc, np.array(lum[:], dtype=np.int32).reshape(8, 8), np.array(chr_[:], dtype=np.int32).reshape(8, 8)


def block_counts(w, h, ratio):
    n = (C.c_long * 3)()
    rc = lib().orc_block_counts(w, h, ratio, n)
    if rc:
        raise ValueError("bad dims %dx%d ratio %d" % (w, h, ratio))
    return [int(n[0]), int(n[1]), int(n[2])]


def coefficients(r, g, b, w, h, quality, ratio=0, oob=None, want_dct=False, threads=1):
    r, g, b = _planes(r, g, b)
    n = block_counts(w, h, ratio)
    zz = np.zeros((sum(n), 64), dtype=np.int16)
    dct = np.zeros((sum(n), 64), dtype=np.float64) if want_dct else None
    oobp = None
    if oob is not None:
        oob = np.ascontiguousarray(oob, dtype=np.uint8).reshape(24)
        oobp = _u8(oob)
    lib().orc_set_threads(threads)
    rc = lib().orc_coefficients(_u8(r), _u8(g), _u8(b), w, h, quality, ratio, oobp, zz.ctypes.data,
                                dct.ctypes.data if want_dct else None)
    lib().orc_set_threads(1)
    if rc:
        raise ValueError("orc_coefficients rc=%d" % rc)
    return (zz, dct) if want_dct else zz


def set_standard(on):
    """standards-conforming arithmetic (textbook Cb, no last-column shift) on/off"""
    lib().orc_set_standard(1 if on else 0)


def pad_planes(r, g, b, w, h, ratio):
    """edge replication to whole MCUs, as jpgb200_encode_planar_any does on the device"""
    mw, mh = (16 if ratio else 8), (16 if ratio == 2 else 8)
    pw, ph = (w + mw - 1) // mw * mw, (h + mh - 1) // mh * mh
    out = [np.pad(np.asarray(p, np.uint8).reshape(h, w), ((0, ph - h), (0, pw - w)), mode="edge").reshape(-1) for p in (r, g, b)]
    return out[0], out[1], out[2], pw, ph


def dc_sequence(zz_comp, mode):
    zz_comp = np.ascontiguousarray(zz_comp, dtype=np.int16)
    n = zz_comp.shape[0]
    d = np.zeros(n, dtype=np.int32)
    lib().orc_dc_sequence(zz_comp.ctypes.data, n, mode, d.ctypes.data)
    return d


def histograms(zz, n, dc_mode, sym_mode):
    zz = np.ascontiguousarray(zz, dtype=np.int16)
    nn = (C.c_long * 3)(*n)
    f = np.zeros((4, 257), dtype=np.int32)
    lib().orc_histograms(zz.ctypes.data, nn, dc_mode, sym_mode, f.ctypes.data)
    return f


def huffman_table(t):
    code = np.zeros(256, dtype=np.uint16)
    ln = np.zeros(256, dtype=np.uint8)
    lib().orc_huffman_table(t, code.ctypes.data, ln.ctypes.data)
    return code, ln


def encode_scan(zz, w, h, ratio):
    zz = np.ascontiguousarray(zz, dtype=np.int16)
    cap = zz.size * 4 + 1024
    out = np.zeros(cap, dtype=np.uint8)
    tb = C.c_uint64(0)
    n = lib().orc_encode_scan(zz.ctypes.data, w, h, ratio, out.ctypes.data, cap, C.byref(tb))
    assert n > 0
    return out[:n].copy(), int(tb.value)


def encode_scan_raw(zz, w, h, ratio):
    zz = np.ascontiguousarray(zz, dtype=np.int16)
    cap = zz.size * 4 + 1024
    out = np.zeros(cap, dtype=np.uint8)
    tb = C.c_uint64(0)
    n = lib().orc_encode_scan_raw(zz.ctypes.data, w, h, ratio, out.ctypes.data, cap, C.byref(tb))
    assert n > 0
    return out[:n].copy(), int(tb.value)


def block_bits(zz, w, h, ratio):
    zz = np.ascontiguousarray(zz, dtype=np.int16)
    bits = np.zeros(zz.shape[0], dtype=np.uint32)
    lib().orc_block_bits(zz.ctypes.data, w, h, ratio, bits.ctypes.data)
    return bits


def jfif_header(w, h, quality, ratio):
    out = np.zeros(1024, dtype=np.uint8)
    n = lib().orc_jfif_header(w, h, quality, ratio, out.ctypes.data, 1024)
    assert n > 0
    return out[:n].copy()


class HuffSpecs(C.Structure):
    """OrcHuffSpecs: BITS/HUFFVAL of the four tables (0 DC lum, 1 DC chr, 2 AC lum, 3 AC chr)."""
    _fields_ = [("nr", (C.c_uint8 * 17) * 4), ("vals", (C.c_uint8 * 256) * 4), ("nvals", C.c_int * 4)]

    def lengths(self, t):
        """code length per symbol of table t (0 = absent)"""
        ln = np.zeros(256, dtype=np.int32)
        k = 0
        for l in range(1, 17):
            for _ in range(self.nr[t][l]):
                ln[self.vals[t][k]] = l
                k += 1
        return ln


def specs_standard():
    sp = HuffSpecs()
    lib().orc_specs_standard(C.byref(sp))
    return sp


def specs_optimal(freq):
    """T.81 K.2 tables from SPEC-mode histograms freq[4][257] (lum_DC, lum_AC, chrom_DC, chrom_AC)."""
    f = np.ascontiguousarray(freq, dtype=np.int32)
    sp = HuffSpecs()
    lib().orc_specs_optimal(f.ctypes.data, C.byref(sp))
    return sp


def encode_jfif_specs(zz, w, h, quality, ratio, sp):
    """header (with sp's DHT) + scan coded with sp + EOI"""
    zz = np.ascontiguousarray(zz, dtype=np.int16)
    hdr = np.zeros(1024, dtype=np.uint8)
    n = lib().orc_jfif_header_specs(w, h, quality, ratio, C.byref(sp), hdr.ctypes.data, 1024)
    assert n > 0
    cap = zz.size * 4 + 1024
    out = np.zeros(cap, dtype=np.uint8)
    m = lib().orc_encode_scan_specs(zz.ctypes.data, w, h, ratio, C.byref(sp), out.ctypes.data, cap, None)
    assert m > 0
    return np.concatenate([hdr[:n], out[:m], np.array([0xFF, 0xD9], np.uint8)])


def encode_jfif(r, g, b, w, h, quality, ratio=0, oob=None):
    r, g, b = _planes(r, g, b)
    cap = w * h * 6 + 4096
    out = np.zeros(cap, dtype=np.uint8)
    oobp = None
    if oob is not None:
        oob = np.ascontiguousarray(oob, dtype=np.uint8).reshape(24)
        oobp = _u8(oob)
    n = lib().orc_encode_jfif(_u8(r), _u8(g), _u8(b), w, h, quality, ratio, oobp, out.ctypes.data, cap)
    assert n > 0
    return out[:n].copy()


def bmp_load(path):
    u8p = C.POINTER(C.c_uint8)
    r, g, b = u8p(), u8p(), u8p()
    w, h = C.c_int(), C.c_int()
    rc = lib().orc_bmp_load(path.encode(), C.byref(r), C.byref(g), C.byref(b), C.byref(w), C.byref(h))
    if rc:
        raise IOError("orc_bmp_load rc=%d" % rc)
    n = w.value * h.value
    out = [np.ctypeslib.as_array(p, shape=(n,)).copy() for p in (r, g, b)]
    for p in (r, g, b):
        lib().orc_free(p)
    return out[0], out[1], out[2], w.value, h.value


def synth_plane(w, h, seed, frame, channel, kind):
    a = np.zeros((h, w), dtype=np.uint8)
    lib().orc_synth_plane(a.ctypes.data, w, h, seed, frame, channel, kind)
    return a


def synth_planes(w, h, seed, frame=0, kind=0):
    return tuple(synth_plane(w, h, seed, frame, c, kind) for c in range(3))


def time_encode(r, g, b, w, h, quality, ratio=0, threads=1):
    r, g, b = _planes(r, g, b)
    lib().orc_set_threads(threads)
    t = lib().orc_time_encode(_u8(r), _u8(g), _u8(b), w, h, quality, ratio)
    lib().orc_set_threads(1)
    return t


def write_bmp(path, r, g, b, w, h, trailing=0):
    """Standard 24-bit bottom-up BMP whose planes, read through the REFERENCE loader
    (src/bitmap.c:129-135), come back as exactly (r, g, b): the loader labels file
    bytes 0,1,2 of each pixel red, green, blue and takes rows from the end of the file."""
    assert (w * 3) % 4 == 0
    r, g, b = [np.asarray(p, dtype=np.uint8).reshape(h, w) for p in (r, g, b)]
    hdr = bytearray(54)
    hdr[0:2] = b"BM"
    size = 54 + w * h * 3 + trailing
    hdr[2:6] = size.to_bytes(4, "little")
    hdr[10:14] = (54).to_bytes(4, "little")
    hdr[14:18] = (40).to_bytes(4, "little")
    hdr[18:22] = w.to_bytes(4, "little")
    hdr[22:26] = h.to_bytes(4, "little")
    hdr[26:28] = (1).to_bytes(2, "little")
    hdr[28:30] = (24).to_bytes(2, "little")
    hdr[34:38] = (w * h * 3).to_bytes(4, "little")
    px = np.stack([r, g, b], axis=-1)  # file byte order = loader's (red, green, blue)
    body = px[::-1].tobytes()  # loader row i (1-based from top) = file row from the end
    with open(path, "wb") as f:
        f.write(bytes(hdr) + body + b"\0" * trailing)


def read_ref_dump(path):
    """Parse a dump written by oracle/_ref/ref_harness (layout in oracle/ref_harness.c)."""
    raw = np.fromfile(path, dtype=np.uint8)
    hdr = raw[:20].view(np.int32)
    assert hdr[0] == 0x4A524546
    w, h, q, nb = int(hdr[1]), int(hdr[2]), int(hdr[3]), int(hdr[4])
    p = 20
    oob = raw[p:p + 24].copy()
    p += 24
    npx = w * h
    planes = [raw[p + i * npx:p + (i + 1) * npx].copy() for i in range(3)]
    p += 3 * npx
    zz = raw[p:p + 3 * nb * 64 * 4].view(np.int32).reshape(3 * nb, 64).copy()
    p += 3 * nb * 64 * 4
    dp = raw[p:p + 3 * nb * 4].view(np.int32).reshape(3, nb).copy()
    p += 3 * nb * 4
    freq = raw[p:p + 4 * 257 * 4].view(np.int32).reshape(4, 257).copy()
    return dict(w=w, h=h, quality=q, nblocks=nb, oob=oob, r=planes[0], g=planes[1], b=planes[2], zz=zz,
                dpcm=dp, freq=freq)


def run_ref_dump(bmp_path, quality, out_path, harness=None):
    subprocess.check_call([harness or REF_HARNESS, "dump", bmp_path, str(quality), out_path],
                          stdout=subprocess.DEVNULL)
    return read_ref_dump(out_path)
