"""BASELINE.json configs[4] at full size: one 32768x32768 4:2:2 q75 image, MCU-row partitioned over the
box's GPUs (torchrun); timing of mgpu.encode_partitioned and identity with a single-GPU encode."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root: run as `torchrun ... profiles/run_config5_partition.py W H Q RATIO`
sys.path.insert(0, ROOT)

def main():
    import torch, torch.distributed as dist
    import __graft_entry__ as ge
    w, h, q, ratio = [int(x) for x in sys.argv[1:5]]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pkg = ge.load_package()
    enc = pkg.Encoder(local)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); enc.set_stream(stream.cuda_stream)
    planes = torch.empty((3, h, w), dtype=torch.uint8, device=dev)
    enc.synth_device(planes.data_ptr(), w, h, 1004, 0, 0)
    torch.cuda.synchronize()
    row0, bh = pkg.mgpu.plan_bands(h, world)[rank]
    ptrs = tuple(planes[c].data_ptr() + row0 * w for c in range(3))
    header = torch.from_numpy(pkg.binding.jfif_header(w, h, q, ratio)).to(dev)
    segment = torch.empty(w * bh * 2 + 65536, dtype=torch.uint8, device=dev)
    out = torch.empty(w * h * 2 + 65536, dtype=torch.uint8, device=dev) if rank == 0 else None
    times = []
    for it in range(4):
        dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        n = pkg.mgpu.encode_partitioned(enc, ptrs, w, bh, q, ratio, header, segment, out, dist, dev)
        torch.cuda.synchronize(); dist.barrier(); times.append(time.perf_counter() - t0)
    if rank == 0:
        cap = w * h * 2 + 65536
        one = torch.empty(cap, dtype=torch.uint8, device=dev); ln = torch.zeros(1, dtype=torch.int64, device=dev)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for it in range(2):
            ev0.record(stream)
            enc.encode_device(planes[0].data_ptr(), planes[1].data_ptr(), planes[2].data_ptr(), w, h, q, ratio, one.data_ptr(), cap, ln.data_ptr())
            ev1.record(stream); torch.cuda.synchronize()
        n1 = int(ln.item())
        same = (n == n1) and bool((out[:n] == one[:n1]).all().item())
        best = min(times[1:])
        print("CONFIG5 world=%d %dx%d q%d ratio%d bytes=%d identical=%s partitioned_best_s=%.4f (%.0f Mpixel/s) single_gpu_ms=%.2f (%.0f Mpixel/s)"
              % (world, w, h, q, ratio, n, same, best, w * h / best / 1e6, ev0.elapsed_time(ev1), w * h / ev0.elapsed_time(ev1) / 1e3), flush=True)
    dist.barrier(); dist.destroy_process_group()

if __name__ == "__main__":
    main()
