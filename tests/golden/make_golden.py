"""Regenerate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/ref_harness).

Run in the build container (needs /root/reference to have been compiled by
`make -C oracle ref`).  Each vector stores what the reference's own stage
functions produced: quantised zig-zag coefficients (before dpcm), the dpcm()
output, the run|size histograms, the planes its BMP loader returned (as a
sha256) and the 24 out-of-bounds bytes block W/8-1 read.
cam.bmp / tiger.bmp are the reference's own images (src/images/), copied here
as input fixtures; synthetic inputs are regenerated from (seed, kind).
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402

CASES = [
    # name, source, quality
    ("cam_q50", ("bmp", "cam.bmp"), 50),
    ("tiger_q50", ("bmp", "tiger.bmp"), 50),
    ("tiger_q75", ("bmp", "tiger.bmp"), 75),
    ("syn_natural_128x64_q75", ("syn", 128, 64, 1001, 0, 0), 75),
    ("syn_noise_256x128_q90", ("syn", 256, 128, 1002, 1, 0), 90),
    ("syn_natural_64x128_q30_trail2", ("syn", 64, 128, 1003, 0, 2), 30),
    ("syn_noise_320x240_q97", ("syn", 320, 240, 1004, 1, 0), 97),
    ("syn_natural_96x48_q24", ("syn", 96, 48, 1005, 0, 0), 24),
    ("syn_natural_512x256_q75", ("syn", 512, 256, 1006, 0, 0), 75),
]


def main():
    tmp = "/tmp/golden_tmp"
    os.makedirs(tmp, exist_ok=True)
    for name, src, q in CASES:
        if src[0] == "bmp":
            path = os.path.join(HERE, src[1])
            meta = dict(kind="bmp", file=src[1])
        else:
            _, w, h, seed, kind, trail = src
            r, g, b = O.synth_planes(w, h, seed, 0, kind)
            path = os.path.join(tmp, name + ".bmp")
            O.write_bmp(path, r, g, b, w, h, trailing=trail)
            meta = dict(kind="syn", w=w, h=h, seed=seed, synth_kind=kind, trailing=trail)
        d = O.run_ref_dump(path, q, os.path.join(tmp, name + ".bin"))
        assert np.abs(d["zz"]).max() < 32768
        planes_sha = hashlib.sha256(d["r"].tobytes() + d["g"].tobytes() + d["b"].tobytes()).hexdigest()
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            w=d["w"], h=d["h"], quality=q, oob=d["oob"], zz=d["zz"].astype(np.int16), dpcm=d["dpcm"],
            freq=d["freq"], planes_sha256=planes_sha, meta=repr(meta),
        )
        print(name, d["w"], d["h"], q, "nnz", int((d["zz"] != 0).sum()), planes_sha[:12])
    # the reference's only known-answer vector: stdout of test_dct() (src/jpg_driver.c:54-150)
    import subprocess
    out = os.path.join(HERE, "test_dct_stdout.txt")
    subprocess.check_call([O.REF_HARNESS, "testdct", out])
    print("wrote", out)


if __name__ == "__main__":
    main()
