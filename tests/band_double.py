"""TEST DOUBLE for the MCU-row band protocol (CPU, gloo tests only): answers
band_transform / band_bits / band_pack / band_finish from the oracle's whole-image stream,
and asserts that everything the orchestration hands it (DC predictors, bit phase, carry byte)
equals the truth.  Never used by the product path."""
import ctypes

import numpy as np

import oracle_lib as O


class OracleBandEngine:
    def __init__(self, planes, width, height, quality, ratio, row0, band_h):
        self.w, self.h, self.ratio = width, height, ratio
        r, g, b = planes
        self.zz = O.coefficients(r, g, b, width, height, quality, ratio)
        self.n = O.block_counts(width, height, ratio)
        self.bits = O.block_bits(self.zz, width, height, ratio).astype(np.int64)
        self.raw, self.total_bits = O.encode_scan_raw(self.zz, width, height, ratio)
        bpm = 3 if ratio == 0 else 4
        mcus_per_row = (width // 8) if ratio == 0 else (width // 16)
        self.s_lo = (row0 // 8) * mcus_per_row * bpm
        self.s_hi = ((row0 + band_h) // 8) * mcus_per_row * bpm
        self.m_lo, self.m_hi = self.s_lo // bpm, self.s_hi // bpm
        self.bit_base = int(self.bits[: self.s_lo].sum())
        self.band_bits_true = int(self.bits[self.s_lo:self.s_hi].sum())

    def _last_dc_before(self, m):
        """last DC of every component among MCUs [0, m)"""
        if m == 0:
            return [0, 0, 0]
        n, ratio = self.n, self.ratio
        y_last = m - 1 if ratio == 0 else 2 * m - 1
        return [int(self.zz[y_last, 0]), int(self.zz[n[0] + m - 1, 0]), int(self.zz[n[0] + n[1] + m - 1, 0])]

    def band_transform(self, d_r, d_g, d_b, w, band_h, has_halo, quality, ratio):
        assert has_halo == (self.m_lo > 0)
        return self._last_dc_before(self.m_hi)

    def band_bits(self, pred_dc):
        assert list(pred_dc) == self._last_dc_before(self.m_lo), "wrong DC predictor handed to the band"
        return self.band_bits_true

    def band_pack(self, bit_phase, is_last):
        assert bit_phase == (self.bit_base & 7), "wrong bit phase"
        self.is_last = is_last
        end = self.bit_base + self.band_bits_true
        assert is_last == (end == self.total_bits)
        owned = ((self.bit_base & 7) + self.band_bits_true) >> 3
        if is_last and (end & 7):
            owned += 1
        tail = 0
        if not is_last and (end & 7):
            tail = int(self.raw[end >> 3]) & (0xFF << (8 - (end & 7))) & 0xFF
        self.owned = owned
        return owned, tail

    def band_finish(self, carry_byte, segment_ptr, cap):
        b0 = self.bit_base >> 3
        ph = self.bit_base & 7
        if ph:
            want = int(self.raw[b0]) & (0xFF << (8 - ph)) & 0xFF
            assert carry_byte == want, "wrong carry byte"
        data = self.raw[b0:b0 + self.owned].copy()
        end = self.bit_base + self.band_bits_true
        if self.is_last and (end & 7):
            data[-1] |= (1 << (8 - (end & 7))) - 1  # 1-padding
        out = bytearray()
        for v in data.tolist():
            out.append(v)
            if v == 0xFF:
                out.append(0)
        assert len(out) <= cap
        ctypes.memmove(segment_ptr, bytes(out), len(out))
        return len(out)
