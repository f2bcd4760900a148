"""torchrun worker for the multi-GPU tests (launched by tests/test_gpu_multi.py and profiles/run_config5.sh):
   python -m torch.distributed.run --nproc-per-node N tests/partition_worker.py W H Q RATIO OUT.npy [reps]
Every rank generates ONLY its band of the synthetic image (plus the halo row above it) on its own GPU and
runs
  * jpgb200_mgpu_encode (C entry: two ncclAllGathers, KS storing straight into rank 0's buffer), and
  * mgpu.encode_partitioned (the earlier torch.distributed orchestration),
rank 0 compares both with a single-GPU encode of the whole image (built there from the same generator),
saves the stitched stream for the parent test (which decodes it and checks a band against the oracle) and
prints PARTITION ... identical=True plus device timings of the C path."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SEED = 1005


def main():
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge

    w, h, q, ratio = [int(x) for x in sys.argv[1:5]]
    out_path = sys.argv[5]
    reps = int(sys.argv[6]) if len(sys.argv) > 6 else 1
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pkg = ge.load_package()
    enc = pkg.Encoder(local)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    enc.set_stream(stream.cuda_stream)
    row0, bh = pkg.mgpu.plan_bands(h, world)[rank]
    halo = 1 if rank > 0 else 0
    rows = bh + halo
    band = torch.empty((3, rows, w), dtype=torch.uint8, device=dev)  # this rank's band + the row above it, nothing else
    enc.synth_device_rows(band.data_ptr(), w, rows, row0 - halo, SEED, 0, 0)
    torch.cuda.synchronize()
    ptrs = tuple(band[c].data_ptr() + halo * w for c in range(3))
    cap = w * h * 2 + 65536

    # ---- the C entry over NCCL + peer memory
    uid = torch.from_numpy(pkg.binding.nccl_unique_id() if rank == 0 else np.zeros(128, np.uint8)).to(dev)
    dist.broadcast(uid, 0)
    mg = pkg.binding.MgpuEncoder(enc, uid.cpu().numpy(), rank, world, cap)
    mg.encode(ptrs[0], ptrs[1], ptrs[2], w, bh, h, q, ratio)
    n_c = mg.result()
    times = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        e0.record(stream)
        mg.encode(ptrs[0], ptrs[1], ptrs[2], w, bh, h, q, ratio)
        e1.record(stream)
        n_c = mg.result()
        tt = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        times.append(float(tt.item()))
    c_stream = None
    if rank == 0:
        c_stream = enc.d2h(mg.output_ptr(), n_c).copy()

    # ---- the torch.distributed orchestration (kept as a cross-check)
    header = torch.from_numpy(pkg.binding.jfif_header(w, h, q, ratio)).to(dev)
    segment = torch.empty(w * bh * 2 + 65536, dtype=torch.uint8, device=dev)
    out = torch.empty(cap, dtype=torch.uint8, device=dev) if rank == 0 else None
    n_py = pkg.mgpu.encode_partitioned(enc, ptrs, w, bh, q, ratio, header, segment, out, dist, dev)
    torch.cuda.synchronize()
    ok = True
    if rank == 0:
        py_stream = out[:n_py].cpu().numpy()
        # single-GPU encode of the same image for comparison (rank 0 builds the whole image for that)
        planes = torch.empty((3, h, w), dtype=torch.uint8, device=dev)
        enc.synth_device(planes.data_ptr(), w, h, SEED, 0, 0)
        one = torch.empty(cap, dtype=torch.uint8, device=dev)
        ln = torch.zeros(1, dtype=torch.int64, device=dev)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(2):
            ev0.record(stream)
            enc.encode_device(planes[0].data_ptr(), planes[1].data_ptr(), planes[2].data_ptr(), w, h, q, ratio,
                              one.data_ptr(), cap, ln.data_ptr())
            ev1.record(stream)
            torch.cuda.synchronize()
        n1 = int(ln.item())
        single = one[:n1].cpu().numpy()
        same_c = (n_c == n1) and bool((c_stream == single).all())
        same_py = (n_py == n1) and bool((py_stream == single).all())
        ok = same_c and same_py
        np.save(out_path, c_stream if n_c < (64 << 20) else c_stream[: (64 << 20)])
        best = min(times)
        print("PARTITION world=%d %dx%d q%d ratio%d bytes=%d single=%d identical=%s (c_nccl=%s torch_dist=%s) "
              "c_path_best_ms=%.3f mpixel_s=%.0f single_gpu_ms=%.3f single_mpixel_s=%.0f"
              % (world, w, h, q, ratio, n_c, n1, ok, same_c, same_py, best * 1e3, w * h / best / 1e6, ev0.elapsed_time(ev1),
                 w * h / ev0.elapsed_time(ev1) / 1e3), flush=True)
    dist.barrier()
    mg.close()
    dist.destroy_process_group()
    if not ok:
        sys.exit(3)


if __name__ == "__main__":
    main()
