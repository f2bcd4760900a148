"""ctypes binding of libjpgb200.so (include/jpgb200.h) for tests/ and bench.py.

Python is plumbing only: every call goes through the C ABI a C host would use.  If the
shared library is missing or no CUDA device is usable this module raises -- there is no
CPU fallback.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libjpgb200.so")

DC_REF, DC_SPEC = 0, 1
SYM_REF, SYM_SPEC = 0, 1
NUM_KERNELS = 6  # JPGB200_NUM_KERNELS

ERR_NO_DEVICE, ERR_BAD_DIMS, ERR_BAD_QUALITY, ERR_CUDA, ERR_OUT_TOO_SMALL, ERR_BAD_ARG, ERR_NOMEM = -1, -2, -3, -4, -5, -6, -7

_lib = None

# name -> (restype, argtypes); also the list tests check against include/jpgb200.h
_u8p = C.POINTER(C.c_uint8)
_vp = C.c_void_p
_SIGS = {
    "jpgb200_create": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "jpgb200_destroy": (None, [_vp]),
    "jpgb200_strerror": (C.c_char_p, [C.c_int]),
    "jpgb200_last_error": (C.c_char_p, [_vp]),
    "jpgb200_set_stream": (C.c_int, [_vp, _vp]),
    "jpgb200_sync": (C.c_int, [_vp]),
    "jpgb200_set_oob_bytes": (C.c_int, [_vp, _vp]),
    "jpgb200_set_staged_entropy": (C.c_int, [_vp, C.c_int]),
    "jpgb200_set_entropy_form": (C.c_int, [_vp, C.c_int]),
    "jpgb200_entropy_form": (C.c_int, [_vp]),
    "jpgb200_set_plain_transform": (C.c_int, [_vp, C.c_int]),
    "jpgb200_set_optimized_huffman": (C.c_int, [_vp, C.c_int]),
    "jpgb200_set_standard_mode": (C.c_int, [_vp, C.c_int]),
    "jpgb200_set_batch_overlap": (C.c_int, [_vp, C.c_int]),
    "jpgb200_set_fused_transform": (C.c_int, [_vp, C.c_int]),
    "jpgb200_set_fused_tap": (C.c_int, [_vp, C.c_int]),
    "jpgb200_set_frame_groups": (C.c_int, [_vp, C.c_int]),
    "jpgb200_has_ab_forms": (C.c_int, []),
    "jpgb200_set_pipelined_batch": (C.c_int, [_vp, C.c_int]),
    "jpgb200_last_coefficients": (C.c_int, [_vp, _vp]),
    "jpgb200_huffman_specs": (C.c_int, [_vp, _vp, _vp, _vp]),
    "jpgb200_encode_planar": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "jpgb200_encode_planar_any": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "jpgb200_encode_rgb": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "jpgb200_encode_device": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_size_t, _vp]),
    "jpgb200_encode_batch_device": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_size_t, _vp]),
    "jpgb200_encode_batch_host": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_size_t, _vp]),
    "jpgb200_band_transform": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "jpgb200_band_bits": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_ulonglong)]),
    "jpgb200_band_pack": (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(C.c_ulonglong), C.POINTER(C.c_uint8)]),
    "jpgb200_band_finish": (C.c_int, [_vp, C.c_int, _vp, C.c_size_t, C.POINTER(C.c_ulonglong)]),
    "jpgb200_jfif_header": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "jpgb200_coefficients": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "jpgb200_dc_sequence": (C.c_int, [_vp, C.c_int, _vp]),
    "jpgb200_histograms": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "jpgb200_block_bit_offsets": (C.c_int, [_vp, _vp]),
    "jpgb200_num_blocks": (C.c_longlong, [C.c_int, C.c_int, C.c_int, _vp]),
    "jpgb200_stage_dct": (C.c_int, [_vp, _vp, C.c_longlong]),
    "jpgb200_stage_quantise": (C.c_int, [_vp, _vp, C.c_longlong, _vp]),
    "jpgb200_stage_zigzag": (C.c_int, [_vp, _vp, C.c_longlong, _vp]),
    "jpgb200_quant_tables": (C.c_int, [C.c_int, _vp, _vp]),
    "jpgb200_stage_preprocess": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    "jpgb200_stage_subsample422": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int]),
    "jpgb200_stage_dpcm": (C.c_int, [_vp, _vp, C.c_longlong]),
    "jpgb200_stage_histograms": (C.c_int, [_vp, _vp, C.c_longlong, C.c_int, _vp, _vp]),
    "jpgb200_encode_coefficients": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "jpgb200_device_count": (C.c_int, []),
    "jpgb200_malloc": (C.c_int, [_vp, C.POINTER(_vp), C.c_size_t]),
    "jpgb200_free": (C.c_int, [_vp, _vp]),
    "jpgb200_nccl_unique_id": (C.c_int, [_vp]),
    "jpgb200_nccl_init": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, _vp, C.c_int]),
    "jpgb200_nccl_destroy": (C.c_int, [_vp]),
    "jpgb200_mgpu_create": (C.c_int, [C.POINTER(_vp), _vp, _vp, C.c_int, C.c_int, C.c_size_t]),
    "jpgb200_mgpu_destroy": (C.c_int, [_vp]),
    "jpgb200_mgpu_encode": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "jpgb200_mgpu_result": (C.c_int, [_vp, C.POINTER(C.c_ulonglong)]),
    "jpgb200_mgpu_output": (_vp, [_vp]),
    "jpgb200_host_alloc": (C.c_int, [C.POINTER(_vp), C.c_size_t]),
    "jpgb200_host_alloc_ex": (C.c_int, [C.POINTER(_vp), C.c_size_t, C.c_int]),
    "jpgb200_device_numa": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "jpgb200_copy_probe": (C.c_int, [_vp, _vp, C.c_int, C.c_size_t, _vp, C.c_size_t]),
    "jpgb200_host_free": (C.c_int, [_vp]),
    "jpgb200_memcpy_h2d": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "jpgb200_memcpy_d2h": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "jpgb200_synth_device": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_uint64, C.c_uint32, C.c_int]),
    "jpgb200_synth_device_rows": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint32, C.c_int]),
    "jpgb200_set_profiling": (C.c_int, [_vp, C.c_int]),
    "jpgb200_kernel_times": (C.c_int, [_vp, _vp, _vp]),
    "jpgb200_reset_profile": (C.c_int, [_vp]),
    "jpgb200_launch_count": (C.c_longlong, [_vp]),
    "jpgb200_kernel_name": (C.c_char_p, [C.c_int]),
}


def exported_names():
    return sorted(_SIGS)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s is missing: run `python jpeg-encoder-and-decoder_b200/build.py` "
                               "(the encode path has no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class Jpgb200Error(RuntimeError):
    def __init__(self, code, detail=""):
        self.code = code
        msg = lib().jpgb200_strerror(code).decode()
        super().__init__("jpgb200 error %d: %s %s" % (code, msg, detail))


def num_blocks(w, h, ratio):
    n = (C.c_longlong * 3)()
    t = lib().jpgb200_num_blocks(w, h, ratio, n)
    if t < 0:
        raise Jpgb200Error(int(t))
    return [int(n[0]), int(n[1]), int(n[2])]


def quant_tables(quality):
    lum = np.zeros(64, np.int32)
    chr_ = np.zeros(64, np.int32)
    rc = lib().jpgb200_quant_tables(quality, lum.ctypes.data, chr_.ctypes.data)
    return rc, lum.reshape(8, 8), chr_.reshape(8, 8)


def jfif_header(w, h, quality, ratio):
    out = np.zeros(1024, np.uint8)
    n = C.c_size_t(0)
    rc = lib().jpgb200_jfif_header(w, h, quality, ratio, out.ctypes.data, 1024, C.byref(n))
    if rc:
        raise Jpgb200Error(rc)
    return out[: n.value].copy()


def _c(a, dtype=np.uint8):
    return np.ascontiguousarray(a, dtype=dtype).reshape(-1)


class Encoder:
    """One context on one GPU (jpgb200_create / jpgb200_destroy)."""

    def __init__(self, device=0):
        self._h = _vp()
        rc = lib().jpgb200_create(C.byref(self._h), device)
        if rc:
            raise Jpgb200Error(rc)
        self.device = device

    def close(self):
        if self._h:
            lib().jpgb200_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise Jpgb200Error(rc, lib().jpgb200_last_error(self._h).decode())

    # ---- configuration
    def set_stream(self, cuda_stream_ptr):
        self._ck(lib().jpgb200_set_stream(self._h, cuda_stream_ptr))

    def sync(self):
        self._ck(lib().jpgb200_sync(self._h))

    def set_staged_entropy(self, on):
        self._ck(lib().jpgb200_set_staged_entropy(self._h, 1 if on else 0))

    def set_entropy_form(self, mode):
        """KE's form: 0 automatic (follows the content), 1 tight (7 CTAs per SM at 4:4:4), 2 roomy"""
        self._ck(lib().jpgb200_set_entropy_form(self._h, int(mode)))

    def entropy_form(self):
        return int(lib().jpgb200_entropy_form(self._h))

    def set_optimized_huffman(self, on):
        self._ck(lib().jpgb200_set_optimized_huffman(self._h, 1 if on else 0))

    def huffman_specs(self):
        """(nr[4][17], vals[4][256], nvals[4]) of the tables the last encode used"""
        nr = np.zeros((4, 17), dtype=np.uint8)
        vals = np.zeros((4, 256), dtype=np.uint8)
        nv = np.zeros(4, dtype=np.int32)
        self._ck(lib().jpgb200_huffman_specs(self._h, nr.ctypes.data, vals.ctypes.data, nv.ctypes.data))
        return nr, vals, nv

    def set_plain_transform(self, on):
        self._ck(lib().jpgb200_set_plain_transform(self._h, int(on)))

    def set_oob_bytes(self, oob):
        oob = _c(oob)
        assert oob.size == 24
        self._ck(lib().jpgb200_set_oob_bytes(self._h, oob.ctypes.data))

    # ---- encode
    def encode_planar(self, r, g, b, w, h, quality, ratio=0, cap=None):
        r, g, b = _c(r), _c(g), _c(b)
        cap = cap or (w * h * 3 + 65536)
        out = np.empty(cap, np.uint8)
        n = C.c_size_t(0)
        self._ck(lib().jpgb200_encode_planar(self._h, r.ctypes.data, g.ctypes.data, b.ctypes.data, w, h, quality, ratio,
                                            out.ctypes.data, cap, C.byref(n)))
        return out[: n.value].copy()

    def encode_planar_any(self, r, g, b, w, h, quality, ratio=0, cap=None):
        """any width/height: edge replication to whole MCUs on the device, SOF0 with the true size"""
        r, g, b = _c(r), _c(g), _c(b)
        cap = cap or ((w + 16) * (h + 8) * 3 + 65536)
        out = np.empty(cap, np.uint8)
        n = C.c_size_t(0)
        self._ck(lib().jpgb200_encode_planar_any(self._h, r.ctypes.data, g.ctypes.data, b.ctypes.data, w, h, quality, ratio,
                                                out.ctypes.data, cap, C.byref(n)))
        return out[: n.value].copy()

    def set_batch_overlap(self, on):
        self._ck(lib().jpgb200_set_batch_overlap(self._h, 1 if on else 0))

    def set_fused_transform(self, on):
        """K1E (transform + chunk encode in one kernel; opt-in, measured slower than K1 -> KE) on/off"""
        self._ck(lib().jpgb200_set_fused_transform(self._h, 1 if on else 0))

    def set_pipelined_batch(self, on):
        """K1 of frame f+1 beside KE of frame f (three streams, two coefficient buffers)"""
        self._ck(lib().jpgb200_set_pipelined_batch(self._h, 1 if on else 0))

    def set_frame_groups(self, on):
        """multi-frame launches for batches of small frames (default on)"""
        self._ck(lib().jpgb200_set_frame_groups(self._h, 1 if on else 0))

    def set_fused_tap(self, on):
        """K1E additionally writes its coefficients to HBM (read them back with last_coefficients)"""
        self._ck(lib().jpgb200_set_fused_tap(self._h, 1 if on else 0))

    def set_standard_mode(self, on):
        self._ck(lib().jpgb200_set_standard_mode(self._h, 1 if on else 0))

    def encode_rgb(self, rgb, w, h, quality, ratio=0, cap=None):
        rgb = _c(rgb)
        cap = cap or (w * h * 3 + 65536)
        out = np.empty(cap, np.uint8)
        n = C.c_size_t(0)
        self._ck(lib().jpgb200_encode_rgb(self._h, rgb.ctypes.data, w, h, quality, ratio, out.ctypes.data, cap, C.byref(n)))
        return out[: n.value].copy()

    def encode_device(self, d_r, d_g, d_b, w, h, quality, ratio, d_out, cap, d_len=None):
        self._ck(lib().jpgb200_encode_device(self._h, d_r, d_g, d_b, w, h, quality, ratio, d_out, cap, d_len))

    def encode_batch_device(self, d_planes, n_frames, w, h, quality, ratio, d_out, out_stride, d_lens=None):
        self._ck(lib().jpgb200_encode_batch_device(self._h, d_planes, n_frames, w, h, quality, ratio, d_out, out_stride, d_lens))

    def encode_batch_host(self, frame_ptrs, w, h, quality, ratio, out_ptrs, cap):
        n = len(frame_ptrs)
        fa = (_vp * n)(*frame_ptrs)
        oa = (_vp * n)(*out_ptrs)
        lens = (C.c_size_t * n)()
        self._ck(lib().jpgb200_encode_batch_host(self._h, fa, n, w, h, quality, ratio, oa, cap, lens))
        return [int(x) for x in lens]

    def copy_probe(self, frame_ptrs, in_bytes, out_ptrs, out_bytes):
        """the copies of encode_batch_host without its kernels (the PCIe ceiling of an end-to-end encode)"""
        n = len(frame_ptrs)
        fa = (_vp * n)(*frame_ptrs)
        oa = (_vp * n)(*out_ptrs)
        self._ck(lib().jpgb200_copy_probe(self._h, fa, n, in_bytes, oa, out_bytes))

    # ---- MCU-row bands (one band per rank; see mgpu.py)
    def band_transform(self, d_r, d_g, d_b, w, band_h, has_halo, quality, ratio):
        dc = (C.c_int * 3)()
        self._ck(lib().jpgb200_band_transform(self._h, d_r, d_g, d_b, w, band_h, 1 if has_halo else 0, quality, ratio, dc))
        return [int(dc[0]), int(dc[1]), int(dc[2])]

    def band_bits(self, pred_dc):
        bits = C.c_ulonglong(0)
        pd = (C.c_int * 3)(*pred_dc)
        self._ck(lib().jpgb200_band_bits(self._h, pd, C.byref(bits)))
        return int(bits.value)

    def band_pack(self, bit_phase, is_last):
        own = C.c_ulonglong(0)
        tail = C.c_uint8(0)
        self._ck(lib().jpgb200_band_pack(self._h, bit_phase, 1 if is_last else 0, C.byref(own), C.byref(tail)))
        return int(own.value), int(tail.value)

    def band_finish(self, carry_byte, d_segment, cap):
        n = C.c_ulonglong(0)
        self._ck(lib().jpgb200_band_finish(self._h, carry_byte, d_segment, cap, C.byref(n)))
        return int(n.value)

    # ---- parity taps
    def coefficients(self, r, g, b, w, h, quality, ratio=0):
        r, g, b = _c(r), _c(g), _c(b)
        n = num_blocks(w, h, ratio)
        zz = np.empty((sum(n), 64), np.int16)
        self._ck(lib().jpgb200_coefficients(self._h, r.ctypes.data, g.ctypes.data, b.ctypes.data, w, h, quality, ratio, zz.ctypes.data))
        self._n = n
        return zz

    def last_coefficients(self, w, h, ratio=0):
        nb = num_blocks(w, h, ratio)
        zz = np.empty((sum(nb), 64), dtype=np.int16)
        self._ck(lib().jpgb200_last_coefficients(self._h, zz.ctypes.data))
        return zz

    def dc_sequence(self, mode):
        d = np.empty(sum(self._n), np.int32)
        self._ck(lib().jpgb200_dc_sequence(self._h, mode, d.ctypes.data))
        return d

    def histograms(self, dc_mode, sym_mode):
        f = np.zeros((4, 257), np.int32)
        self._ck(lib().jpgb200_histograms(self._h, dc_mode, sym_mode, f.ctypes.data))
        return f

    def block_bit_offsets(self):
        off = np.empty(sum(self._n) + 1, np.uint64)
        self._ck(lib().jpgb200_block_bit_offsets(self._h, off.ctypes.data))
        return off

    # ---- staged FP64 block API
    def stage_dct(self, blocks):
        b = np.ascontiguousarray(blocks, np.float64).reshape(-1, 64).copy()
        self._ck(lib().jpgb200_stage_dct(self._h, b.ctypes.data, b.shape[0]))
        return b

    def stage_quantise(self, blocks, table):
        b = np.ascontiguousarray(blocks, np.float64).reshape(-1, 64).copy()
        t = _c(table, np.int32)
        self._ck(lib().jpgb200_stage_quantise(self._h, b.ctypes.data, b.shape[0], t.ctypes.data))
        return b

    def stage_zigzag(self, blocks):
        b = np.ascontiguousarray(blocks, np.float64).reshape(-1, 64)
        zz = np.zeros((b.shape[0], 64), np.int32)
        self._ck(lib().jpgb200_stage_zigzag(self._h, b.ctypes.data, b.shape[0], zz.ctypes.data))
        return zz

    def stage_preprocess(self, r, g, b, w, h):
        r, g, b = _c(r), _c(g), _c(b)
        n = (w // 8) * (h // 8)
        Y, Cb, Cr = (np.zeros((n, 64), np.float64) for _ in range(3))
        self._ck(lib().jpgb200_stage_preprocess(self._h, r.ctypes.data, g.ctypes.data, b.ctypes.data, w, h,
                                               Y.ctypes.data, Cb.ctypes.data, Cr.ctypes.data))
        return Y, Cb, Cr

    def stage_subsample422(self, blocks, bw, bh):
        blk = np.ascontiguousarray(blocks, np.float64).reshape(-1, 64)
        out = np.zeros((blk.shape[0] // 2, 64), np.float64)
        self._ck(lib().jpgb200_stage_subsample422(self._h, blk.ctypes.data, out.ctypes.data, bw, bh))
        return out

    def stage_dpcm(self, dc):
        d = np.ascontiguousarray(dc, np.int32).copy()
        self._ck(lib().jpgb200_stage_dpcm(self._h, d.ctypes.data, d.size))
        return d

    def stage_histograms(self, zz, sym_mode, freq_dc=None, freq_ac=None):
        z = np.ascontiguousarray(zz, np.int32).reshape(-1, 64)
        fd = np.zeros(257, np.int32) if freq_dc is None else np.ascontiguousarray(freq_dc, np.int32).copy()
        fa = np.zeros(257, np.int32) if freq_ac is None else np.ascontiguousarray(freq_ac, np.int32).copy()
        self._ck(lib().jpgb200_stage_histograms(self._h, z.ctypes.data, z.shape[0], sym_mode, fd.ctypes.data, fa.ctypes.data))
        return fd, fa

    def encode_coefficients(self, zz, w, h, quality, ratio=0, cap=None):
        z = np.ascontiguousarray(zz, np.int16)
        cap = cap or (w * h * 3 + 65536)
        out = np.empty(cap, np.uint8)
        n = C.c_size_t(0)
        self._ck(lib().jpgb200_encode_coefficients(self._h, z.ctypes.data, w, h, quality, ratio, out.ctypes.data, cap, C.byref(n)))
        return out[: n.value].copy()

    # ---- device helpers
    def malloc(self, nbytes):
        p = _vp()
        self._ck(lib().jpgb200_malloc(self._h, C.byref(p), nbytes))
        return p.value

    def free(self, p):
        self._ck(lib().jpgb200_free(self._h, p))

    def h2d(self, dptr, arr):
        arr = np.ascontiguousarray(arr)
        self._ck(lib().jpgb200_memcpy_h2d(self._h, dptr, arr.ctypes.data, arr.nbytes))

    def d2h(self, dptr, nbytes):
        out = np.empty(nbytes, np.uint8)
        self._ck(lib().jpgb200_memcpy_d2h(self._h, out.ctypes.data, dptr, nbytes))
        return out

    def synth_device(self, d_planes, w, h, seed, frame=0, kind=0):
        self._ck(lib().jpgb200_synth_device(self._h, d_planes, w, h, seed, frame, kind))

    def synth_device_rows(self, d_planes, w, rows, row0, seed, frame=0, kind=0):
        self._ck(lib().jpgb200_synth_device_rows(self._h, d_planes, w, rows, row0, seed, frame, kind))

    # ---- measurement
    def set_profiling(self, on):
        self._ck(lib().jpgb200_set_profiling(self._h, int(on)))

    def reset_profile(self):
        self._ck(lib().jpgb200_reset_profile(self._h))

    def kernel_times(self):
        ms = (C.c_double * NUM_KERNELS)()
        ln = (C.c_longlong * NUM_KERNELS)()
        self._ck(lib().jpgb200_kernel_times(self._h, ms, ln))
        names = [lib().jpgb200_kernel_name(i).decode() for i in range(NUM_KERNELS)]
        return {names[i]: (float(ms[i]), int(ln[i])) for i in range(NUM_KERNELS)}

    def launch_count(self):
        return int(lib().jpgb200_launch_count(self._h))


def nccl_unique_id():
    """128 bytes of ncclGetUniqueId (rank 0 makes it, the caller broadcasts it)"""
    buf = np.zeros(128, np.uint8)
    rc = lib().jpgb200_nccl_unique_id(buf.ctypes.data)
    if rc:
        raise Jpgb200Error(rc)
    return buf


class MgpuEncoder:
    """jpgb200_mgpu_*: one band of a large image per rank, stitched in rank 0's buffer (include/jpgb200.h)."""

    def __init__(self, enc, unique_id, rank, world, out_cap):
        self.enc, self.rank, self.world = enc, rank, world
        self._comm = _vp()
        uid = np.ascontiguousarray(unique_id, dtype=np.uint8)
        rc = lib().jpgb200_nccl_init(C.byref(self._comm), world, rank, uid.ctypes.data, enc.device)
        if rc:
            raise Jpgb200Error(rc)
        self._mg = _vp()
        rc = lib().jpgb200_mgpu_create(C.byref(self._mg), enc._h, self._comm, rank, world, out_cap)
        if rc:
            raise Jpgb200Error(rc, lib().jpgb200_last_error(enc._h).decode())

    def encode(self, d_r, d_g, d_b, w, band_h, image_h, quality, ratio):
        rc = lib().jpgb200_mgpu_encode(self._mg, d_r, d_g, d_b, w, band_h, image_h, quality, ratio)
        if rc:
            raise Jpgb200Error(rc, lib().jpgb200_last_error(self.enc._h).decode())

    def result(self):
        n = C.c_ulonglong(0)
        rc = lib().jpgb200_mgpu_result(self._mg, C.byref(n))
        if rc:
            raise Jpgb200Error(rc, lib().jpgb200_last_error(self.enc._h).decode())
        return int(n.value)

    def output_ptr(self):
        return lib().jpgb200_mgpu_output(self._mg)

    def close(self):
        if self._mg:
            lib().jpgb200_mgpu_destroy(self._mg)
            self._mg = None
        if self._comm:
            lib().jpgb200_nccl_destroy(self._comm)
            self._comm = None


def has_ab_forms():
    return bool(lib().jpgb200_has_ab_forms())


def device_numa(device=0):
    """(NUMA node of the GPU's PCIe function or -1, NUMA nodes visible to this process)"""
    a, b = C.c_int(-1), C.c_int(0)
    rc = lib().jpgb200_device_numa(device, C.byref(a), C.byref(b))
    if rc:
        raise Jpgb200Error(rc)
    return int(a.value), int(b.value)


def pinned_empty(nbytes, write_combined=False):
    """numpy view of pinned host memory (jpgb200_host_alloc_ex); keep the returned owner alive."""
    p = _vp()
    rc = lib().jpgb200_host_alloc_ex(C.byref(p), nbytes, 1 if write_combined else 0)
    if rc:
        raise Jpgb200Error(rc)
    buf = (C.c_uint8 * nbytes).from_address(p.value)
    arr = np.frombuffer(buf, dtype=np.uint8)
    return arr, p.value


def pinned_free(ptr):
    lib().jpgb200_host_free(ptr)
