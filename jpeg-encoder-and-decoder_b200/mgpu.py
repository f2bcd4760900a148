"""Multi-GPU orchestration of the encode path; one process per GPU.

Two shardings, both over `torch.distributed` (NCCL on GPUs, gloo in the CPU tests):

* batches of images -- frames are independent: rank r encodes frames r, r+G, r+2G, ...
  (`shard_frames`), no collective on the data path.

* one very large image -- `encode_partitioned`: the image is cut into bands of whole MCU
  rows (`plan_bands`), one band per rank.  The entropy-coded stream is a single bit string
  with DC prediction running across bands, so four tiny allgathers stitch it:
      0. {last DC of Y/Cb/Cr}       -> every band's DC predictor (it changes the bit length
                                       of the band's first blocks, so it must come first)
      1. {band bits}                -> every band's bit offset in the stream
      2. {trailing partial byte}    -> the leading bits of the next band's first byte
      3. {stuffed segment length}   -> every band's byte offset in the final stream
  A band owns the stream bytes whose LAST bit lies in its bit range, so the stuffed
  segments are byte-aligned; ranks > 0 then send their segment straight into its place in
  rank 0's output buffer (NCCL send/recv = a peer-to-peer copy over NVLink).

The per-band work is done by an "engine" with band_transform / band_bits / band_pack /
band_finish (binding.Encoder on a GPU; the C entry points are jpgb200_band_*).  Python is
plumbing only.
"""


def shard_frames(n_frames, rank, world):
    """frame indices of this rank: f mod world == rank"""
    return list(range(rank, n_frames, world))


def plan_bands(height, world):
    """[(row0, band_height)] per rank: whole MCU rows (8 pixel rows), as even as possible"""
    if height % 8:
        raise ValueError("height must be a multiple of 8")
    rows = height // 8
    if rows < world:
        raise ValueError("fewer MCU rows (%d) than ranks (%d)" % (rows, world))
    base, rem = divmod(rows, world)
    bands, r0 = [], 0
    for k in range(world):
        n = base + (1 if k < rem else 0)
        bands.append((r0 * 8, n * 8))
        r0 += n
    return bands


def _allgather_i64(dist, values, device, group=None):
    import torch

    world = dist.get_world_size(group)
    t = torch.tensor(values, dtype=torch.int64, device=device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    return [[int(x) for x in o.cpu().tolist()] for o in out]


def encode_partitioned(engine, band_ptrs, width, band_height, quality, ratio, header, segment, out, dist, device,
                       group=None):
    """Encode this rank's band and stitch the JFIF stream on rank 0.

    engine    : band_transform(d_r, d_g, d_b, w, band_h, has_halo, q, ratio) -> last_dc[3]
                band_bits(pred_dc) -> bits
                band_pack(bit_phase, is_last) -> (owned_bytes, tail_byte)
                band_finish(carry_byte, segment_ptr, cap) -> segment_len
    band_ptrs : (r, g, b) addresses of the band's first row (row above valid when rank > 0)
    header    : uint8 tensor, the JFIF marker segments of the WHOLE image (jpgb200_jfif_header)
    segment   : uint8 tensor on `device` receiving this band's stuffed bytes
    out       : uint8 tensor on `device`, rank 0's output buffer (other ranks: None)
    returns   : total stream length on rank 0, this rank's segment length elsewhere
    """
    rank, world = dist.get_rank(group), dist.get_world_size(group)

    def grank(k):  # send/recv take GLOBAL ranks also when `group` is a sub-group
        return dist.get_global_rank(group, k) if group is not None and hasattr(dist, "get_global_rank") else k

    # Every exchange also carries a status word, so that a rank whose engine call failed cannot leave its peers
    # blocked in a collective or a recv: everybody sees the failure in the same allgather and raises.
    err = None

    def step(fn, fallback):
        nonlocal err
        if err is not None:
            return fallback
        try:
            return fn()
        except Exception as exc:  # noqa: BLE001 -- reported to every rank below
            err = exc
            return fallback

    def exchange(values):
        got = _allgather_i64(dist, [0 if err is None else 1] + list(values), device, group)
        bad = [k for k, g in enumerate(got) if g[0]]
        if bad:
            raise RuntimeError("encode_partitioned failed on rank(s) %s: %s" % (bad, err if err is not None else "see that rank"))
        return [g[1:] for g in got]

    last_dc = step(lambda: engine.band_transform(band_ptrs[0], band_ptrs[1], band_ptrs[2], width, band_height, rank > 0, quality,
                                                 ratio), [0, 0, 0])
    g0 = exchange(list(last_dc))                                      # exchange 0: DC predictors
    pred = g0[rank - 1] if rank > 0 else [0, 0, 0]
    bits = step(lambda: engine.band_bits(pred), 0)
    g1 = exchange([bits])                                             # exchange 1: bit offsets
    bit_base = sum(g[0] for g in g1[:rank])
    owned, tail = step(lambda: engine.band_pack(bit_base & 7, rank == world - 1), (0, 0))
    g2 = exchange([tail])                                             # exchange 2: boundary bytes
    carry = g2[rank - 1][0] if rank > 0 else 0
    seg_len = step(lambda: engine.band_finish(carry, segment.data_ptr(), segment.numel()), 0)
    hl = header.numel()
    cap0 = out.numel() if rank == 0 else 0
    g3 = exchange([seg_len, cap0])                                    # exchange 3: byte offsets (+ rank 0's capacity)
    lens = [g[0] for g in g3]
    total = hl + sum(lens) + 2
    if total > g3[0][1]:  # checked on EVERY rank before any point-to-point call
        raise ValueError("output buffer too small: need %d bytes" % total)
    if rank == 0:
        out[:hl].copy_(header)
        off = hl
        out[off:off + lens[0]].copy_(segment[:lens[0]])
        off += lens[0]
        for k in range(1, world):
            if lens[k]:
                dist.recv(out[off:off + lens[k]], src=grank(k), group=group)
            off += lens[k]
        out[off] = 0xFF
        out[off + 1] = 0xD9
        return total
    if seg_len:
        dist.send(segment[:seg_len], dst=grank(0), group=group)
    return seg_len
