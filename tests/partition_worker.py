"""torchrun worker for the multi-GPU tests (launched by tests/test_gpu_multi.py):
   python -m torch.distributed.run --nproc-per-node N tests/partition_worker.py W H Q RATIO OUT.npy
Every rank generates its band of the synthetic image on its own GPU (with the halo row above
it), runs mgpu.encode_partitioned over NCCL, and rank 0 saves the stitched stream."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge

    w, h, q, ratio = [int(x) for x in sys.argv[1:5]]
    out_path = sys.argv[5]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pkg = ge.load_package()
    enc = pkg.Encoder(local)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    enc.set_stream(stream.cuda_stream)
    # the whole synthetic image on every GPU (test sizes are small enough); a rank only READS its band + halo
    planes = torch.empty((3, h, w), dtype=torch.uint8, device=dev)
    enc.synth_device(planes.data_ptr(), w, h, 1005, 0, 0)
    torch.cuda.synchronize()
    row0, bh = pkg.mgpu.plan_bands(h, world)[rank]
    ptrs = tuple(planes[c].data_ptr() + row0 * w for c in range(3))
    header = torch.from_numpy(pkg.binding.jfif_header(w, h, q, ratio)).to(dev)
    segment = torch.empty(w * bh * 2 + 65536, dtype=torch.uint8, device=dev)
    out = torch.empty(w * h * 2 + 65536, dtype=torch.uint8, device=dev) if rank == 0 else None
    n = pkg.mgpu.encode_partitioned(enc, ptrs, w, bh, q, ratio, header, segment, out, dist, dev)
    torch.cuda.synchronize()
    if rank == 0:
        # single-GPU encode of the same image for comparison
        cap = w * h * 2 + 65536
        one = torch.empty(cap, dtype=torch.uint8, device=dev)
        ln = torch.zeros(1, dtype=torch.int64, device=dev)
        enc.encode_device(planes[0].data_ptr(), planes[1].data_ptr(), planes[2].data_ptr(), w, h, q, ratio,
                          one.data_ptr(), cap, ln.data_ptr())
        torch.cuda.synchronize()
        n1 = int(ln.item())
        same = (n == n1) and bool((out[:n] == one[:n1]).all().item())
        np.save(out_path, out[:n].cpu().numpy() if n < (64 << 20) else out[: (64 << 20)].cpu().numpy())
        print("PARTITION world=%d %dx%d q%d ratio%d bytes=%d single=%d identical=%s" % (world, w, h, q, ratio, n, n1, same), flush=True)
        if not same:
            dist.barrier()
            dist.destroy_process_group()
            sys.exit(3)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
