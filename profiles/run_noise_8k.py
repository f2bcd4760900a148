import sys, io, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import __graft_entry__ as ge
pkg = ge.load_package()
enc = pkg.Encoder(0)
enc.set_pipelined_batch(0)
enc.set_batch_overlap(0)  # isolated kernel spans
W, H = 7680, 4320
planes = torch.empty((2, 3, H, W), dtype=torch.uint8, device="cuda")
out = torch.empty((2, 3 * W * H), dtype=torch.uint8, device="cuda")
lens = torch.zeros(2, dtype=torch.int64, device="cuda")
from PIL import Image
Image.MAX_IMAGE_PIXELS = None
for kind, q in ((1, 75), (1, 97), (0, 75)):
    for f in range(2):
        enc.synth_device(planes.data_ptr() + f * 3 * W * H, W, H, 1003, f, kind)
    torch.cuda.synchronize()
    enc.set_profiling(1)
    for it in range(2):
        enc.reset_profile()
        enc.encode_batch_device(planes.data_ptr(), 2, W, H, q, 0, out.data_ptr(), 3 * W * H, lens.data_ptr())
        torch.cuda.synchronize()
    kt = enc.kernel_times()
    enc.set_profiling(0)
    l = lens.cpu().numpy()
    jf = out[0, : int(l[0])].cpu().numpy()
    im = Image.open(io.BytesIO(jf.tobytes())); im.load()
    print("kind", kind, "q", q, "bytes/px %.3f" % (l[0] / (W * H)), {k: round(v[0] / max(v[1], 1) * 1e3, 1) for k, v in kt.items()}, im.size)
