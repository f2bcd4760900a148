import sys
sys.path.insert(0, ".")
import torch
import __graft_entry__ as ge
pkg = ge.load_package(); enc = pkg.Encoder(0)
for (W, H, q, ratio, name) in [(3840, 2160, 90, 1, "config3 4K 4:2:2 q90"), (1920, 1080, 75, 0, "config2 1080p 4:4:4 q75"), (3840, 2160, 75, 2, "4K 4:2:0 q75 (16-row padded: 2176)")]:
    if ratio == 2: H = 2176
    F = 64
    planes = torch.empty((F, 3, H, W), dtype=torch.uint8, device="cuda")
    out = torch.empty((F, 2 * W * H), dtype=torch.uint8, device="cuda"); lens = torch.zeros(F, dtype=torch.int64, device="cuda")
    for f in range(F): enc.synth_device(planes.data_ptr() + f * 3 * W * H, W, H, 1002, f, 0)
    torch.cuda.synchronize()
    s = torch.cuda.Stream(); torch.cuda.set_stream(s); enc.set_stream(s.cuda_stream)
    for it in range(3): enc.encode_batch_device(planes.data_ptr(), F, W, H, q, ratio, out.data_ptr(), 2 * W * H, lens.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for it in range(5): enc.encode_batch_device(planes.data_ptr(), F, W, H, q, ratio, out.data_ptr(), 2 * W * H, lens.data_ptr())
    e1.record(s); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5 / F
    print("%s: %.1f us/frame, %.0f Mpixel/s, %.3f B/px" % (name, ms * 1e3, W * H / ms / 1e3, float(lens.float().mean()) / (W * H)), flush=True)
