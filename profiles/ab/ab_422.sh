# usage: ab_422.sh lib1 lib2 ... : 4K 4:2:2 q90 batch of 64 frames, per-kernel times
cp jpeg-encoder-and-decoder_b200/libjpgb200.so /tmp/default.so
for l in "$@"; do
  if [ "$l" = default ]; then cp /tmp/default.so jpeg-encoder-and-decoder_b200/libjpgb200.so; else cp profiles/ab/libs/$l.so jpeg-encoder-and-decoder_b200/libjpgb200.so; fi
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config3 or synthetic_vs_oracle" 2>&1 | tail -1
  python - <<PY
import sys
sys.path.insert(0, ".")
import torch
import __graft_entry__ as ge
pkg = ge.load_package(); enc = pkg.Encoder(0)
W, H, q, ratio, F = 3840, 2160, 90, 1, 64
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
enc.reset_profile(); enc.set_profiling(1); enc.set_pipelined_batch(False); enc.set_batch_overlap(False)
enc.encode_batch_device(planes.data_ptr(), F, W, H, q, ratio, out.data_ptr(), 2 * W * H, lens.data_ptr()); torch.cuda.synchronize()
kt = enc.kernel_times()
print("$l 4K 4:2:2 q90: %.2f us/frame, %.0f Mpixel/s" % (ms * 1e3, W * H / ms / 1e3), {k: round(v[0] / F * 1e3, 2) for k, v in kt.items() if v[1]}, flush=True)
PY
done
cp /tmp/default.so jpeg-encoder-and-decoder_b200/libjpgb200.so
