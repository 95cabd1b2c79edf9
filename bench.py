#!/usr/bin/env python
"""bench.py — 1080p H.264 encode throughput of the B200 macroblock pipeline (BASELINE.json metric).

A "step" = one batch-frame: every one of the S independent 1080p streams resident on this GPU advances by
one picture (mode decision + ME + transform/quant + reconstruction + deblocking + border expansion on the
GPU, CAVLC on host threads).  Workload = BASELINE.json configs[2]: synthetic 1920x1080 I420, constant
QP 26, camera mode, single slice, complexity HIGH ("full ME": SATD costs, all partitions), IDR + P...;
the bitstream is bit-identical to the reference encoder's (tests/test_gpu_encoder.py).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--streams S] [--impl reference]

N > 1: one process per GPU under torch.distributed (launched by the driver with torchrun); streams are
independent, there is no data-path collective ("scaling": "weak"); NCCL is used only for the barrier and the
max-over-ranks of the timed interval.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

W, H, QP, FPS = 1920, 1080, 26, 30.0
CLIP_FRAMES = 16                      # distinct synthetic pictures; played forward/backward (no scene cuts)
ALG_BYTES_PER_MB = 2016               # cur 384 + ref 384 + recon 384 + levels 768 + MVs 64 + meta 32
MBS_PER_FRAME = 120 * 68


def clip_order(n):
    seq = list(range(CLIP_FRAMES)) + list(range(CLIP_FRAMES - 2, 0, -1))
    return [seq[i % len(seq)] for i in range(n)]


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows if len(r) > 2 + i)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons}


def make_clip():
    import h264lib
    return h264lib.synth_clip(W, H, CLIP_FRAMES)


# ---------------------------------------------------------------------------------------------------------------
def reference_arm(args, rank, world):
    """The reference's own CPU encoder (oracle/_ref, unmodified, public API) on the box's host cores: one
    single-threaded encoder instance per host thread, each coding its own 1080p stream — the same batched
    independent-stream workload as the GPU arm.  A step = `frames_per_step` pictures per stream."""
    if rank != 0:
        return
    import h264lib
    if not h264lib.have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built on this machine"}))
        return
    R = C.CDLL(h264lib.REFSHIM_SO)
    R.ref_enc_open.restype = C.c_void_p
    R.ref_enc_open.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float]
    R.ref_enc_frames.restype = C.c_long
    R.ref_enc_frames.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong)]
    R.ref_enc_close.argtypes = [C.c_void_p]
    clip = make_clip()
    fsz = W * H * 3 // 2
    cores = min(os.cpu_count() or 1, args.ref_threads or (os.cpu_count() or 1))
    fps_guess = 25.0
    frames_per_step = 1
    order = clip_order((args.steps + args.warmup) * frames_per_step)
    encs = [R.ref_enc_open(W, H, QP, 2, 1, FPS) for _ in range(cores)]
    ts = [C.c_longlong(0) for _ in range(cores)]

    def run_steps(lo, hi):
        def worker(i):
            for st in range(lo, hi):
                f = order[(st + 3 * i) % len(order)]
                R.ref_enc_frames(encs[i], clip[f * fsz:].ctypes.data, W, H, 1, C.byref(ts[i]))
        th = [threading.Thread(target=worker, args=(i,)) for i in range(cores)]
        [t.start() for t in th]
        [t.join() for t in th]

    run_steps(0, args.warmup)
    t0 = time.perf_counter()
    run_steps(args.warmup, args.warmup + args.steps)
    dt = time.perf_counter() - t0
    for e in encs:
        R.ref_enc_close(e)
    fps = cores * args.steps * frames_per_step / dt
    print(json.dumps({
        "impl": "reference", "metric": "1080p_encode_fps", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "encode synthetic 1920x1080 I420, constant QP 26, camera, single slice, complexity HIGH, CAVLC "
                               "(BASELINE.json configs[2]); %d independent streams, 1 picture per stream per step" % cores,
                   "streams": cores},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "reference",
                         "sample": "%d single-thread reference encoders (C-only build, no nasm on this image) x %d pictures each"
                                   % (cores, args.steps)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline_sample(seconds_budget=15.0):
    """reference encoder, 1 thread, bounded sample of the same workload (rank 0, N=1)."""
    import h264lib
    if not h264lib.have_ref():
        return None
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_encoder_golden import ref_encode
    clip = make_clip()
    fsz = W * H * 3 // 2
    n = 24
    order = clip_order(n)
    yuv = np.concatenate([clip[f * fsz:(f + 1) * fsz] for f in order])
    _, _, secs = ref_encode(yuv, W, H, n, QP, FPS, complexity=2, threads=1)
    return {"value": n / secs, "unit": "frames/s", "cores": 1, "kind": "reference",
            "sample": "%d pictures of the bench clip through the reference's ISVCEncoder::EncodeFrame, 1 thread, "
                      "C-only build (USE_ASM=No: nasm absent)" % n}


def ncu_traffic(kernel):
    """per-launch DRAM bytes of `kernel` from the committed ncu capture (profiles/ncu_traffic.json), or None"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(kernel)
    except Exception:
        return None


def mc_sad_roofline(L, local, peak):
    """BASELINE metric 2: the MC+SAD unit (b2h264_k_mc_sad, layer 1) against the HBM roofline.  64 stacked 1080p
    planes (cur + padded ref = 302 MB, larger than L2), one candidate per macroblock; algorithmic bytes per MB =
    256 cur + 256 ref + 4 vector + 4 cost."""
    import torch
    from openh264_b200.binding import check
    S, stride, rows_per = 64, 2048, 1152                      # 1088 + 2 x 32 rows of padding per picture
    g = torch.Generator(device="cuda"); g.manual_seed(264)
    cur = torch.randint(0, 256, (S * rows_per, stride), dtype=torch.uint8, device="cuda", generator=g)
    ref = torch.randint(0, 256, (S * rows_per, stride), dtype=torch.uint8, device="cuda", generator=g)
    mbw, mbh = 120, (S * rows_per) // 16 - 4                  # skip two MB rows at either end: the halo stays inside
    o0 = 32 * stride + 32
    n = mbw * mbh
    out = {}
    st = torch.cuda.current_stream().cuda_stream
    for name, frac in (("integer", False), ("quarter", True)):
        mv = torch.randint(-8, 9, (n, 1, 2), dtype=torch.int16, device="cuda", generator=g) * 4
        if frac:
            mv += torch.randint(0, 4, (n, 1, 2), dtype=torch.int16, device="cuda", generator=g)
        cost = torch.empty((n, 1), dtype=torch.int32, device="cuda")
        run = lambda: check(L.b2h264_k_mc_sad(cur.data_ptr() + o0, stride, ref.data_ptr() + o0, stride, mbw, mbh,
                                              mv.data_ptr(), 1, cost.data_ptr(), st))
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 10
        e0.record()
        for _ in range(iters):
            run()
        e1.record()
        torch.cuda.synchronize()
        sec = e0.elapsed_time(e1) / 1e3 / iters
        gbs = n * 520 / sec / 1e9
        out[name] = {"achieved": gbs, "frac": gbs / peak, "us_per_launch": sec * 1e6, "mb_per_launch": n}
    t = ncu_traffic("k_mc_sad_tma")
    return {"kernel": "k_mc_sad_tma (b2h264_k_mc_sad: 8x4-MB tiles staged by TMA bulk tensor copies)", "bound": "hbm", "unit": "GB/s",
            "traffic": t["dram_bytes"] if t and t.get("mb_per_launch") == n else None,
            "peak": peak, "alg_bytes_per_mb": 520, "candidates_per_mb": 1, "achieved": out["integer"]["achieved"],
            "frac": out["integer"]["frac"], "integer_mv": out["integer"], "quarter_pel_mv": out["quarter"],
            "note": "64 stacked padded 1080p planes per launch (302 MB of pixels, larger than L2)"}


# ---------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--streams", type=int, default=256, help="independent 1080p streams per GPU")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--ref-threads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from openh264_b200.binding import BatchEncoder, lib
    from openh264_b200 import shard
    torch.cuda.set_device(local)
    shard.init("nccl", torch.device("cuda", local))        # one process per GPU; replicas only
    S = args.streams
    L = lib(local)
    clip_h = make_clip()
    fsz = W * H * 3 // 2
    clip_d = torch.from_numpy(clip_h).cuda()
    clip_pinned = torch.from_numpy(clip_h).pin_memory()
    stream = torch.cuda.Stream()
    total = args.steps + args.warmup + 1
    order = clip_order(total + S)

    def run(enc, on_device, n_steps, first):
        """pipelined submit/collect; returns bytes of bitstream produced"""
        nbytes = 0

        def srcs(step):
            base = clip_d.data_ptr() if on_device else None
            out = []
            for s in range(S):
                f = order[step + s % 5]
                out.append(base + f * fsz if on_device else clip_pinned.numpy()[f * fsz:(f + 1) * fsz])
            return out
        enc.submit(srcs(first), on_device=on_device)
        for i in range(1, n_steps + 1):
            if i < n_steps:
                enc.submit(srcs(first + i), on_device=on_device)
            bs, _ = enc.collect()
            nbytes += sum(len(b) for b in bs)
        return nbytes

    results = {}
    launches0 = L.b2h264_launch_count()
    kern_us = []
    d2h_bytes = []
    for mode in ("resident", "e2e"):
        enc = BatchEncoder(W, H, qp=QP, fps=FPS, n_streams=S, device=local,
                           entropy_threads=min(S, max(8, (os.cpu_count() or 8) // world)))
        enc.set_stream(stream.cuda_stream)
        on_dev = mode == "resident"
        run(enc, on_dev, args.warmup, 0)                           # warm-up (includes the IDR pictures)
        torch.cuda.synchronize()
        shard.barrier()
        sampler = ClockSampler(local)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            e0.record(stream)
        t0 = time.perf_counter()
        # timed region: exactly args.steps batch-frames
        nb = 0

        def srcs(step):
            out = []
            for s in range(S):
                f = order[step + s % 5]
                out.append(clip_d.data_ptr() + f * fsz if on_dev else clip_pinned.numpy()[f * fsz:(f + 1) * fsz])
            return out
        first = args.warmup
        enc.submit(srcs(first), on_device=on_dev)
        for i in range(1, args.steps + 1):
            if i < args.steps:
                enc.submit(srcs(first + i), on_device=on_dev)
            bs, _ = enc.collect()
            nb += sum(len(b) for b in bs)
            if mode == "resident":
                kern_us.append(enc.timing_us())
            else:
                d2h_bytes.append(enc.d2h_bytes())
        with torch.cuda.stream(stream):
            e1.record(stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        dev = e0.elapsed_time(e1) / 1e3
        sampler.stop_flag = True
        sampler.join(timeout=2)
        dt, pics = shard.job_totals(max(wall, dev), S * args.steps, device="cuda")   # MAX over ranks / SUM over ranks
        results[mode] = {"dt": dt, "pictures": pics, "wall": wall, "dev": dev, "bytes": nb, "clocks": sampler.summary()}
        enc.close()
    launches = L.b2h264_launch_count() - launches0

    if rank == 0:
        value = results["resident"]["pictures"] / results["resident"]["dt"]
        e2e = results["e2e"]["pictures"] / results["e2e"]["dt"]
        peak, peak_kind = measured_peaks()
        k_enc = float(np.mean([k[0] for k in kern_us])) * 1e-6        # seconds per launch of pad + macroblock wavefront kernel
        k_dbk = float(np.mean([k[1] for k in kern_us])) * 1e-6
        ent = float(np.mean([k[2] for k in kern_us])) * 1e-6
        alg = S * MBS_PER_FRAME * ALG_BYTES_PER_MB
        achieved = alg / k_enc / 1e9
        out = {
            "metric": "1080p_encode_fps", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": results["resident"]["dt"] / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "encode synthetic 1920x1080 I420, constant QP 26, camera, single slice, complexity HIGH, "
                                   "CAVLC (BASELINE.json configs[2]); %d independent streams per GPU, 1 picture per stream "
                                   "per step; bitstream bit-identical to the reference" % S,
                       "streams_per_gpu": S, "parallelism": "replica x%d (independent streams, no collective)" % world,
                       "l2": "inputs larger than L2 (%.0f MB of pictures in flight per step)" % (S * fsz / 1e6)},
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": world * S * fsz, "d2h_bytes_per_step": int(world * np.mean(d2h_bytes)),
                    "d2h_note": "index table + records of the coded macroblocks only, written by the GPU into mapped pinned memory"},
            "gpu_launches": int(launches),
            "clocks": results["resident"]["clocks"],
            "roofline": {"bound": "hbm", "kernel": "k_encode_mbs (macroblock wavefront, all streams)", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (ncu_traffic("k_encode_mbs") or {}).get("dram_bytes") if S == (ncu_traffic("k_encode_mbs") or {}).get("streams") else None,
                         "peak_source": peak_kind,
                         "note": "wavefront kernel is dependency/latency-bound by construction (SURVEY.md §8d); "
                                 "alg bytes = %d B/MB x %d MB/launch" % (ALG_BYTES_PER_MB, S * MBS_PER_FRAME)},
            "breakdown_ms_per_step": {"encode_kernel": k_enc * 1e3, "deblock_expand": k_dbk * 1e3, "host_entropy": ent * 1e3,
                                      "wall_resident": results["resident"]["wall"] / args.steps * 1e3,
                                      "wall_e2e": results["e2e"]["wall"] / args.steps * 1e3},
            "bitstream_kbytes_per_frame": results["resident"]["bytes"] / (S * args.steps) / 1e3,
        }
        if world == 1:
            out["roofline_mc_sad"] = mc_sad_roofline(L, local, peak)
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline_sample()
        print(json.dumps(out))
    if world > 1:
        shard.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
