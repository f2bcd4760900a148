#!/usr/bin/env python
"""bench.py -- encode Mpixel/s of the B200 JPEG encode path on BASELINE.json's headline
configuration: a batch of 8K (7680x4320) RGB frames, 4:4:4, quality 75, one image stream per
GPU (configs[3]); frames are independent, so ranks shard the batch with no collective
("scaling": "weak": every GPU encodes FRAMES frames per step).

A step = one pass of the whole hot path (K1 transform, KE chunk encode, KF 0xFF count, KS stuffed
scatter -> complete JFIF streams) over one batch of synthetic frames.
  value : frames are resident in HBM when the timed region starts (device-resident encode)
  e2e   : the same batch through the C-ABI host-buffer call (jpgb200_encode_batch_host):
          pinned host frames in, JFIF bytes out, H2D and D2H inside the timed region; beside it
          the PCIe ceiling of this host at this N: the same copies with no kernel (jpgb200_copy_probe)
  roofline : the dominant kernel (the longest when each runs alone) against the measured HBM copy
          bandwidth in MEASURED_PEAKS.json.  The timed region is a three-stream pipeline in which K1 and KE
          of neighbouring frames queue for the same SMs, so an event pair on a kernel's stream brackets a
          span that includes that queueing (span_in_timed_region_us); the kernel's DURATION (avg_launch_us,
          frac) comes from CUDA events around every launch of one extra step of the same run with the
          frames' kernels one after the other.  all_timed_kernels holds K1 (north_star's target kernel:
          9 B/pixel) and KE (6 B/pixel + stream); pipeline = (3 B/pixel + stream) / frame time against the
          same peak
  config.strong_scaling / small_frame / worst_case_distribution / config5 : BASELINE.json configs[3] as
          written (256 frames in total), configs[1] (1080p batch), iid-noise frames, configs[4] (one
          32768x32768 image partitioned over the ranks, N > 1 only)
  cpu_baseline : the reference's own C stage functions (oracle/_ref/ref_harness, built from
          the unmodified sources) on a bounded crop of frame 0, one core

`--impl reference` times the reference's CPU implementation on all host cores instead.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

W, H, QUALITY, RATIO = 7680, 4320, 75, 0
NPX = W * H
SEED = 1003  # 1000 + config index
KIND = 0     # "natural"
METRIC = "encode Mpixel/s (8K RGB, 4:4:4 q75)"
UNIT = "Mpixel/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=256, help="frames per GPU per step")
    ap.add_argument("--e2e-frames", type=int, default=64, help="frames per GPU per e2e step")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-worst-case", action="store_true", help="skip the extra noise-distribution measurement")
    ap.add_argument("--no-extras", action="store_true", help="skip the strong-scaling, small-frame and partition legs")
    return ap.parse_args()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clocks and throttle reasons DURING the timed region: an NVML thread sampling every 5 ms (the timed region of a
    fast run is tens of milliseconds: `nvidia-smi -lms 100` misses it), nvidia-smi as the fallback."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    # NVML clocks-event-reason bits
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.nvml = None
        self.samples = []  # (sm MHz, reasons bitmask)
        self.max_mhz = None
        self.stop_flag = False

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
        except Exception:
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)

    def _nvml_loop(self):
        pynvml, h = self.nvml
        while not self.stop_flag:
            try:
                mhz = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                try:
                    rs = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((float(mhz), int(rs)))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        try:
            self.nvml = self._nvml_handle()
            self.max_mhz = float(self.nvml[0].nvmlDeviceGetMaxClockInfo(self.nvml[1], self.nvml[0].NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nvml:
            self.stop_flag = True
            self.thread.join(timeout=1)
            sm = [a for a, _ in self.samples]
            reasons = set()
            for _, rs in self.samples:
                for bit, nm in self.REASONS.items():
                    if rs & bit:
                        reasons.add(nm)
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons),
                    "samples": len(sm), "source": "NVML, one sample per 5 ms of the timed region"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            f = [x.strip() for x in row.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for i, nm in enumerate(names):
                if f[5 + i].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


# ---------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU code (oracle/_ref), all host cores
# ---------------------------------------------------------------------------------------------

def write_crop_bmp(path, w, h, frame, x0=0, y0=0):
    import oracle_lib as O
    import numpy as np

    # identical bytes to the GPU's synthetic frame: generate the crop's pixels of the 8K frame
    full_rows = [O.synth_plane(W, y0 + h, SEED, frame, c, KIND)[y0:y0 + h, x0:x0 + w] for c in range(3)]
    O.write_bmp(path, np.ascontiguousarray(full_rows[0]), np.ascontiguousarray(full_rows[1]),
                np.ascontiguousarray(full_rows[2]), w, h)


def ref_harness_path():
    import oracle_lib as O

    if os.path.exists(O.REF_HARNESS):
        return O.REF_HARNESS
    return None


def time_reference_procs(bmps, quality, harness):
    """one process per BMP (the reference rewrites its global tables: one image per process)"""
    outs = [b + ".json" for b in bmps]
    t0 = time.time()
    procs = [subprocess.Popen([harness, "time", b, str(quality), o], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
             for b, o in zip(bmps, outs)]
    for p in procs:
        p.wait()
    wall = time.time() - t0
    inner = [json.load(open(o))["seconds"] for o in outs]
    return wall, inner


def cpu_baseline_one_core(tmp):
    """bounded sample: a 1920x1080 crop of frame 0 through the unmodified reference, 1 core"""
    harness = ref_harness_path()
    cw, ch = 1920, 1080
    if harness:
        bmp = os.path.join(tmp, "crop.bmp")
        write_crop_bmp(bmp, cw, ch, 0)
        wall, inner = time_reference_procs([bmp], QUALITY, harness)
        out = {"value": cw * ch / inner[0] / 1e6, "unit": UNIT, "cores": 1, "kind": "reference",
               "sample": "1920x1080 crop of frame 0, stages preprocess..dpcm + histogram loops (the reference has no "
                         "bitstream writer), built -O0 -g as shipped (src/Makefile:2), %.1f s" % inner[0]}
        o2 = harness + "_O2"  # the same sources with -O2, for fairness
        if os.path.exists(o2):
            try:
                _, inner2 = time_reference_procs([bmp], QUALITY, o2)
                out["value_O2"] = cw * ch / inner2[0] / 1e6
            except Exception:
                pass
        return out
    import oracle_lib as O
    import numpy as np

    planes = [np.ascontiguousarray(O.synth_plane(W, ch, SEED, 0, c, KIND)[:, :cw]) for c in range(3)]
    t = O.time_encode(planes[0], planes[1], planes[2], cw, ch, QUALITY, RATIO, threads=1)
    return {"value": cw * ch / t / 1e6, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "1920x1080 crop of frame 0 through oracle/jpgb200_oracle.c (FP64 restatement), %.2f s" % t}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_lib as O

    O.lib()
    harness = ref_harness_path()
    cores = os.cpu_count() or 1
    cw, ch = 640, 360  # per-process crop: ~1.7 s of reference time each
    tmp = tempfile.mkdtemp(prefix="jpgb200_ref_")
    bmps = []
    for i in range(cores):
        p = os.path.join(tmp, "crop%d.bmp" % i)
        write_crop_bmp(p, cw, ch, i % 4, x0=(i * cw) % (W - cw))
        bmps.append(p)
    times = []
    kind = "reference" if harness else "port"
    for step in range(args.warmup + args.steps):
        if harness:
            wall, _ = time_reference_procs(bmps, QUALITY, harness)
        else:
            import numpy as np

            t0 = time.time()
            for b in bmps[:1]:
                r, g, bb, w_, h_ = O.bmp_load(b)
                O.coefficients(r, g, bb, w_, h_, QUALITY, RATIO, threads=cores)
            wall = (time.time() - t0)
        if step >= args.warmup:
            times.append(wall)
    npx_step = cw * ch * (cores if harness else 1)
    total = sum(times)
    value = npx_step * len(times) / total / 1e6
    sample = ("%d processes x one %dx%d crop of the 8K synthetic frames per step, unmodified reference stage functions "
              "(preprocess..dpcm + histogram loops; it has no bitstream writer), -O0 -g as shipped" % (cores, cw, ch)) \
        if harness else "oracle port, %d threads, one %dx%d crop per step" % (cores, cw, ch)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "7680x4320 RGB frames, 4:4:4, q75 (bounded crops)", "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------

def run_b200(args):
    import numpy as np
    import torch

    import __graft_entry__ as ge

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the encode path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pkg = ge.load_package()
    enc = pkg.Encoder(local_rank)
    stream = torch.cuda.Stream()  # a real (non-default) stream so torch events bracket our kernels
    torch.cuda.set_stream(stream)
    enc.set_stream(stream.cuda_stream)

    F = args.frames
    cap = NPX  # bytes reserved per output stream (1 B/pixel; natural frames need ~0.3)
    planes = torch.empty((F, 3, H, W), dtype=torch.uint8, device="cuda")
    out = torch.empty((F, cap), dtype=torch.uint8, device="cuda")
    lens = torch.zeros((F,), dtype=torch.int64, device="cuda")
    frame_bytes = 3 * NPX
    for f in range(F):
        enc.synth_device(planes.data_ptr() + f * frame_bytes, W, H, SEED, rank * F + f, KIND)
    torch.cuda.synchronize()

    def step(nf=F):
        enc.encode_batch_device(planes.data_ptr(), nf, W, H, QUALITY, RATIO, out.data_ptr(), cap, lens.data_ptr())

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps):
        """device time of `steps` calls of fn on the bench stream, barrier + synchronize on both sides, max over ranks"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    for _ in range(args.warmup):
        step()
    barrier()
    enc.reset_profile()
    # CUDA events inside the timed region: around every 8th K1 launch and, in other frames, every 8th KE launch
    # (an event record costs ~2.7 us of stream time: bracketing all four kernels of every frame cost 16 us per frame)
    enc.set_profiling(3)
    launches0 = enc.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_total = timed(step, args.steps)
    clocks = sampler.stop()
    headline_form = {1: "tight", 2: "roomy"}.get(enc.entropy_form())
    launches = enc.launch_count() - launches0
    ktimes = enc.kernel_times()
    enc.set_profiling(0)
    # per-kernel breakdown of one more, untimed step (events around every launch; without the
    # batch overlap, under which a tail kernel's event pair also spans its wait behind the next K1)
    enc.reset_profile()
    enc.set_profiling(1)
    enc.set_batch_overlap(False)
    enc.set_pipelined_batch(False)
    step()
    torch.cuda.synchronize()
    kbreak = enc.kernel_times()
    enc.set_profiling(0)
    enc.set_batch_overlap(os.environ.get("JPGB200_OVERLAP", "1") != "0")
    enc.set_pipelined_batch(os.environ.get("JPGB200_PIPELINED", "1") != "0")
    ms_per_step = ms_total / args.steps
    value = world * F * NPX / (ms_per_step * 1e-3) / 1e6
    lens_host = lens.cpu().numpy()
    assert (lens_host > 623).all() and (lens_host <= cap).all(), "an encode failed (length word out of range)"
    bytes_per_px = float(lens_host.mean()) / NPX

    # ---- rooflines from the live event timings: K1 (north_star's target kernel) and the chunk encode
    peak, peak_src = measured_peak()
    traffic_file = {}
    tp = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    if os.path.exists(tp):
        try:
            traffic_file = json.load(open(tp))
        except Exception:
            traffic_file = {}

    def roof(kname, label, alg_bytes, what):
        """One kernel against the roofline.  Its DURATION is measured in this run with CUDA events around every launch
        of one extra step in which the frames' kernels run one after the other (the same kernels, the same inputs, right
        after the timed region).  Inside the timed region itself the batch is pipelined over three streams: an event
        pair on a kernel's stream then brackets a SPAN that includes the time its CTAs queue for SM resources behind
        the other frames' kernels (K1 and KE take turns on the SMs) -- reported beside it, not used for `frac`."""
        ms, n = ktimes.get(kname, (0.0, 0))
        if n == 0:
            return None
        span = ms * 1e-3 / n
        iso_ms = kbreak.get(kname, (0.0, 0))
        iso = iso_ms[0] * 1e-3 / max(iso_ms[1], 1) if iso_ms[1] else None
        dur = iso if iso else span
        # frames per launch: batches run as multi-frame launch groups (one K1 / KE / KF / KS launch per group of frames)
        fpl = max(1, int(round(F / iso_ms[1]))) if iso_ms[1] else 1
        alg_frame = alg_bytes
        alg_bytes = alg_bytes * fpl
        tr = traffic_file.get(kname, {})
        return {"bound": "hbm", "kernel": label, "achieved": alg_bytes / dur / 1e9, "peak": peak, "unit": "GB/s",
                "frac": alg_bytes / dur / 1e9 / peak, "frac_of_nominal_8000": alg_bytes / dur / 1e9 / 8000.0,
                "traffic": (tr["dram_bytes_per_launch"] * fpl / max(1, tr.get("frames_per_launch", 1))) if tr.get("dram_bytes_per_launch") else None,
                "traffic_source": tr.get("source"), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "algorithmic_bytes": what, "avg_launch_us": dur * 1e6,
                "frames_per_launch": fpl, "algorithmic_bytes_per_frame": alg_frame, "us_per_frame": dur * 1e6 / fpl,
                "launches_timed": iso_ms[1] if iso else n,
                "duration_source": ("CUDA events around every launch of one extra step of this run with the frames' kernels one after "
                                    "the other (no other kernel on the GPU), right after the timed region") if iso else
                                   "CUDA events around every 8th launch of the timed region",
                "share_of_step": dur * 1e3 * (F / fpl) / ms_per_step,
                # the event-to-event span of a launch inside the (pipelined) timed region
                "span_in_timed_region_us": span * 1e6, "span_launches_timed": n,
                "frac_of_span": alg_bytes / span / 1e9 / peak,
                "span_what": "every 8th launch of the timed region, CUDA events on the kernel's own stream; in the pipelined batch "
                             "the span of a launch includes the time its CTAs wait for room behind the other frames' kernels",
                "avg_launch_us_isolated": iso * 1e6 if iso else None,
                "frac_isolated": (alg_bytes / iso / 1e9 / peak) if iso else None}

    fused = ktimes.get("k1e_fused", (0.0, 0))[1] > 0
    roofs = {}
    if fused:
        roofs["k1e_fused"] = roof("k1e_fused", "k1e_fused<0>", (3.0 + bytes_per_px) * NPX,
                                  "3 B/px of planes in + the chunk streams out (the coefficients stay on chip)")
    else:
        roofs["k1_transform"] = roof("k1_transform", "k1_transform_pc<0>", 9.0 * NPX,
                                     "9 B/px: 3 B/px of planes in + 6 B/px of int16 coefficients out (SURVEY.md 8(d))")
        roofs["k_encode_chunks"] = roof("k_encode_chunks", "k_encode_chunks<0>", (6.0 + bytes_per_px) * NPX,
                                        "6 B/px of int16 coefficients in + the chunk streams out")
    roofs = {k: v for k, v in roofs.items() if v}
    # the dominant kernel: the longest one when each runs alone
    dominant = max(roofs, key=lambda k: roofs[k]["avg_launch_us_isolated"] or roofs[k]["avg_launch_us"])
    roofline = dict(roofs[dominant])
    roofline["dominant_kernel"] = dominant
    roofline["north_star_target_kernel"] = "k1_transform" if "k1_transform" in roofs else "k1e_fused"
    roofline["all_timed_kernels"] = roofs
    roofline["kernel_ms_per_step_untimed_pass"] = {k: v[0] for k, v in kbreak.items()}
    # the whole pipeline against the same peak: the least it has to move is the planes in and the stream out
    t_frame = ms_per_step * 1e-3 / F
    pipeline = {"algorithmic_bytes_per_frame": (3.0 + bytes_per_px) * NPX, "us_per_frame": t_frame * 1e6,
                "achieved_gbs": (3.0 + bytes_per_px) * NPX / t_frame / 1e9, "frac_of_peak": (3.0 + bytes_per_px) * NPX / t_frame / 1e9 / peak,
                "dram_bytes_per_frame": traffic_file.get("pipeline_dram_bytes_per_frame"),
                "dram_bytes_per_pixel": (traffic_file.get("pipeline_dram_bytes_per_frame") or 0) / NPX or None,
                "traffic_source": traffic_file.get("pipeline_source")}

    # ---- BASELINE.json configs[3] as written: 256 frames in total, sharded over the ranks (strong scaling)
    strong = None
    if not args.no_extras:
        total_frames = 256
        sf = max(1, total_frames // world)
        if sf <= F:
            ms = timed(lambda: step(sf), 2) / 2
            strong = {"scaling": "strong", "frames_total": sf * world, "frames_per_gpu": sf, "ms_per_step": ms,
                      "value": sf * world * NPX / (ms * 1e-3) / 1e6, "unit": UNIT}

    # ---- the other distribution of SURVEY.md 8(d): iid uniform noise, the worst case for the entropy
    # stage and the largest output; a short separate measurement on every rank, reported beside the headline
    noise = None
    try:
        if args.no_worst_case:
            raise RuntimeError("skipped (--no-worst-case)")
        NF = min(F, 8)
        for f in range(NF):
            enc.synth_device(planes.data_ptr() + f * frame_bytes, W, H, SEED, rank * F + f, 1)
        torch.cuda.synchronize()
        nout = torch.empty((NF, 3 * NPX), dtype=torch.uint8, device="cuda")
        nlens = torch.zeros((NF,), dtype=torch.int64, device="cuda")

        def nstep():
            enc.encode_batch_device(planes.data_ptr(), NF, W, H, QUALITY, RATIO, nout.data_ptr(), 3 * NPX, nlens.data_ptr())

        # three warm-up steps, each waited for: KE's form follows the content through a device-side counter that the
        # host polls without blocking, i.e. with the lag of its launch queue (jpgb200_set_entropy_form)
        for _ in range(3):
            nstep()
            torch.cuda.synchronize()
        nms = timed(nstep, 3)
        nl = nlens.cpu().numpy()
        noise = {"distribution": "noise (iid uniform u8)", "value": world * NF * 3 * NPX / (nms * 1e-3) / 1e6, "unit": UNIT,
                 "frames_per_gpu": NF, "steps": 3, "warmup": 3, "entropy_form": {1: "tight", 2: "roomy"}.get(enc.entropy_form()),
                 "compressed_bytes_per_pixel": float(nl.mean()) / NPX, "n_gpus_measured": world,
                 "us_per_frame": nms * 1e3 / (3 * NF), "ratio_to_natural_frame_time": nms * 1e3 / (3 * NF) / (t_frame * 1e6)}
        del nout
        for f in range(NF):  # restore the natural frames for the e2e leg
            enc.synth_device(planes.data_ptr() + f * frame_bytes, W, H, SEED, rank * F + f, KIND)
        torch.cuda.synchronize()
    except Exception as exc:  # the headline must not depend on this extra
        noise = {"error": str(exc)}

    # ---- BASELINE.json configs[1]: small frames (1920x1080 4:4:4 q75), where launches, not bytes, set the pace
    small = None
    if not args.no_extras:
        try:
            small = small_frame_leg(enc, torch, stream, timed, world)
        except Exception as exc:
            small = {"error": str(exc)}

    # ---- e2e: host buffers through the C ABI, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        EF = args.e2e_frames
        ring = 4
        wc = os.environ.get("JPGB200_WC", "0") == "1"
        pins, pouts = [], []
        for i in range(ring):
            a, p = pkg.binding.pinned_empty(frame_bytes, write_combined=wc)
            torch.cuda.synchronize()
            src = planes[i].contiguous().view(-1)
            a[:] = src.cpu().numpy()
            pins.append((a, p))
        for i in range(2 * ring):
            o, p = pkg.binding.pinned_empty(cap)
            pouts.append((o, p))
        enc.set_stream(None)
        in_ptrs = [pins[i % ring][1] for i in range(EF)]
        out_ptrs = [pouts[i % (2 * ring)][1] for i in range(EF)]

        def estep():
            return enc.encode_batch_host(in_ptrs, W, H, QUALITY, RATIO, out_ptrs, cap)

        def host_timed(fn, steps):
            barrier()
            t0 = time.perf_counter()
            r = None
            for _ in range(steps):
                r = fn()
            torch.cuda.synchronize()
            return max_over_ranks(time.perf_counter() - t0), r

        for _ in range(max(1, min(args.warmup, 3))):
            el = estep()
        l0 = enc.launch_count()
        dt, el = host_timed(estep, args.steps)
        launches_e2e = enc.launch_count() - l0
        out_bytes = int(sum(el) // len(el))
        # the ceiling: the same copies (one cudaMemcpyAsync per frame and direction, same streams) with no kernel
        enc.copy_probe(in_ptrs, frame_bytes, out_ptrs, out_bytes)
        dtc, _ = host_timed(lambda: enc.copy_probe(in_ptrs, frame_bytes, out_ptrs, out_bytes), args.steps)
        gpu_node, nodes = pkg.binding.device_numa(local_rank)
        e2e_value = world * EF * NPX * args.steps / dt / 1e6
        ceiling = world * EF * NPX * args.steps / dtc / 1e6
        e2e = {"value": e2e_value, "unit": UNIT,
               "h2d_bytes_per_step": EF * frame_bytes, "d2h_bytes_per_step": int(sum(el)) + 8 * EF,
               "frames_per_step": EF, "ms_per_step": 1e3 * dt / args.steps, "gpu_launches": launches_e2e,
               "timer": "host perf_counter around jpgb200_encode_batch_host, stream-synchronised, max over ranks",
               "pcie_ceiling": ceiling, "frac_of_ceiling": e2e_value / ceiling,
               "pcie_ceiling_what": "jpgb200_copy_probe: the same pinned buffers, sizes and streams, one cudaMemcpyAsync per frame "
                                    "and direction, no kernels, timed the same way at this N",
               "h2d_gbs_per_gpu": EF * frame_bytes * args.steps / dt / 1e9,
               "h2d_gbs_per_gpu_ceiling": EF * frame_bytes * args.steps / dtc / 1e9,
               "pinned": "write-combined" if wc else "cudaHostAlloc default",
               "numa": {"visible_nodes": nodes, "gpu_node": gpu_node, "cpus": len(os.sched_getaffinity(0)),
                        "placement": "single NUMA node visible: nothing to place" if nodes <= 1 else
                                     "pinned buffers first-touched by this rank's thread"}}
        enc.set_stream(stream.cuda_stream)
        for _, p in pins + pouts:
            pkg.binding.pinned_free(p)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        tmp = tempfile.mkdtemp(prefix="jpgb200_cpu_")
        cpu = cpu_baseline_one_core(tmp)

    # ---- BASELINE.json configs[4]: one 32768x32768 image partitioned by MCU rows over the ranks
    config5 = None
    if world > 1 and not args.no_extras:
        try:
            config5 = partition_leg(pkg, enc, torch, dist, rank, world, local_rank)
        except Exception as exc:
            config5 = {"error": str(exc)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "batch of %d synthetic 7680x4320 RGB frames per GPU, 4:4:4, q75 (BASELINE.json configs[3])" % F,
                       "frames_per_gpu_per_step": F, "width": W, "height": H, "quality": QUALITY, "sampling": "4:4:4",
                       "distribution": "natural (16x16 cells + iid noise, SURVEY.md 8(d))", "seed": SEED,
                       "compressed_bytes_per_pixel": bytes_per_px,
                       "l2_policy": "inputs larger than L2: %d distinct frames = %.1f GB per step" % (F, F * frame_bytes / 1e9),
                       "sharding": "frames independent, one image stream per GPU, no collective",
                       "entropy_path": "K1E fused transform + chunk encode" if fused else "K1 -> KE (separate kernels)",
                       "entropy_form": headline_form,
                       "strong_scaling": strong, "worst_case_distribution": noise, "small_frame": small, "config5": config5},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "pipeline": pipeline,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def small_frame_leg(enc, torch, stream, timed, world):
    """configs[1]: a batch of 1920x1080 4:4:4 q75 frames, device resident"""
    w, h, nf = 1920, 1080, 64
    npx = w * h
    cap = npx
    pl = torch.empty((nf, 3, h, w), dtype=torch.uint8, device="cuda")
    out = torch.empty((nf, cap), dtype=torch.uint8, device="cuda")
    lens = torch.zeros((nf,), dtype=torch.int64, device="cuda")
    for f in range(nf):
        enc.synth_device(pl.data_ptr() + f * 3 * npx, w, h, 1001, f, KIND)
    torch.cuda.synchronize()

    def sstep():
        enc.encode_batch_device(pl.data_ptr(), nf, w, h, QUALITY, 0, out.data_ptr(), cap, lens.data_ptr())

    sstep()
    sstep()
    l0 = enc.launch_count()
    ms = timed(sstep, 5) / 5
    launches = (enc.launch_count() - l0) // 5
    ll = lens.cpu().numpy()
    assert (ll > 623).all() and (ll <= cap).all()
    return {"workload": "batch of %d synthetic 1920x1080 RGB frames per GPU, 4:4:4, q75 (BASELINE.json configs[1])" % nf,
            "value": world * nf * npx / (ms * 1e-3) / 1e6, "unit": UNIT, "us_per_frame": ms * 1e3 / nf,
            "launches_per_batch": launches, "compressed_bytes_per_pixel": float(ll.mean()) / npx}


def partition_leg(pkg, enc, torch, dist, rank, world, local_rank):
    """configs[4]: one 32768x32768 4:2:2 q75 image, one band of MCU rows per rank, through the C entry
    jpgb200_mgpu_encode (two ncclAllGathers, segments stored straight into rank 0's buffer over NVLink).
    Every rank generates only its band (+ halo row).  Device time with CUDA events on each rank's stream, max
    over ranks.  Checks (outside the timed region): the stitched stream equals rank 0's single-GPU encode of
    the whole image; the coefficients of a 512-row band equal the oracle's within the coefficient tolerance."""
    import numpy as np

    w = h = 32768
    q, ratio = 75, 1
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.current_stream()
    row0, bh = pkg.mgpu.plan_bands(h, world)[rank]
    halo = 1 if rank > 0 else 0
    rows = bh + halo
    band = torch.empty((3, rows, w), dtype=torch.uint8, device=dev)
    enc.synth_device_rows(band.data_ptr(), w, rows, row0 - halo, 1004, 0, KIND)
    torch.cuda.synchronize()
    ptrs = tuple(band[c].data_ptr() + halo * w for c in range(3))
    cap = w * h // 2 + 65536
    uid = torch.from_numpy(pkg.binding.nccl_unique_id() if rank == 0 else np.zeros(128, np.uint8)).to(dev)
    dist.broadcast(uid, 0)
    mg = pkg.binding.MgpuEncoder(enc, uid.cpu().numpy(), rank, world, cap)
    try:
        mg.encode(ptrs[0], ptrs[1], ptrs[2], w, bh, h, q, ratio)
        n = mg.result()
        best = None
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier()
            torch.cuda.synchronize()
            e0.record(stream)
            mg.encode(ptrs[0], ptrs[1], ptrs[2], w, bh, h, q, ratio)
            e1.record(stream)
            n = mg.result()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = float(t.item()) if best is None else min(best, float(t.item()))
        res = {"workload": "one synthetic 32768x32768 RGB image, 4:2:2, q75, one band of MCU rows per GPU (BASELINE.json configs[4])",
               "path": "jpgb200_mgpu_encode: 2 ncclAllGathers (32 + 24 B per rank) + 8-byte completion allgather, KS stores through a "
                       "CUDA-IPC peer pointer into rank 0's buffer",
               "ms": best, "mpixel_s": w * h / (best * 1e-3) / 1e6, "bytes": n if rank == 0 else None,
               "timer": "CUDA events on each rank's stream around the enqueue of the whole partitioned encode, max over ranks, best of 3"}
        # ---- checks
        same, single_ms = None, None
        if rank == 0:
            planes = torch.empty((3, h, w), dtype=torch.uint8, device=dev)
            enc.synth_device(planes.data_ptr(), w, h, 1004, 0, KIND)
            one = torch.empty(cap, dtype=torch.uint8, device=dev)
            ln = torch.zeros(1, dtype=torch.int64, device=dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(2):
                e0.record(stream)
                enc.encode_device(planes[0].data_ptr(), planes[1].data_ptr(), planes[2].data_ptr(), w, h, q, ratio, one.data_ptr(), cap,
                                  ln.data_ptr())
                e1.record(stream)
                torch.cuda.synchronize()
            n1 = int(ln.item())
            single_ms = e0.elapsed_time(e1)
            stitched = enc.d2h(mg.output_ptr(), n)
            same = (n == n1) and bool((stitched == one[:n1].cpu().numpy()).all())
            del planes, one
        band_ok = None
        if rank == min(1, world - 1):
            import oracle_lib as O  # checker only, outside every timed region

            bh2 = 512
            sub = band[:, halo:halo + bh2, :].cpu().numpy()
            oob = np.zeros(24, np.uint8)
            if halo:
                oob = band[:, 0, w - 8:].cpu().numpy().reshape(-1).astype(np.uint8)
            enc.set_oob_bytes(oob)
            try:
                zz = enc.coefficients(sub[0], sub[1], sub[2], w, bh2, q, ratio)
            finally:
                enc.set_oob_bytes(np.zeros(24, np.uint8))
            ref = O.coefficients(sub[0].reshape(-1), sub[1].reshape(-1), sub[2].reshape(-1), w, bh2, q, ratio, oob=oob,
                                 threads=max(1, (os.cpu_count() or 2) // max(1, world)))
            d = np.abs(zz.astype(np.int32) - ref.astype(np.int32))
            band_ok = bool(d.max() <= 1 and float((d != 0).mean()) <= 1e-3)
        flags = torch.tensor([1.0 if band_ok else 0.0 if band_ok is not None else -1.0], dtype=torch.float64, device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MAX)
        res["identical_to_single_gpu"] = same
        res["single_gpu_ms"] = single_ms
        res["speedup_vs_single_gpu"] = (single_ms / best) if single_ms else None
        res["efficiency_vs_single_gpu"] = (single_ms / best / world) if single_ms else None
        res["bands_equal_oracle"] = bool(flags.item() > 0.5)
        res["bands_equal_oracle_what"] = "quantised coefficients of the first 512 rows of rank %d's band vs oracle/ FP64 (|diff| <= 1 on <= 0.1 %%)" % min(1, world - 1)
        return res
    finally:
        mg.close()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
