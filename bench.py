#!/usr/bin/env python
"""bench.py -- frames/sec of the VCS-h264 interframe hot path on B200 (BASELINE.json metric).

Workload (BASELINE.json configs[1], SURVEY 8d "C2"): synthetic 1080p 8-bit 60-frame clip
(I-P-P-P: 15 I + 45 P), 16x16 macroblocks, +/-16 step-1 full search with the reference's own
cost (wrapped uint8 difference, motion.py:146) and static test (threshold 2000, motion.py:113),
motion-compensated residual, 8x8 DCT, quantise QF=50 (rint -> int8 indices, lossless at this QF), dequantise, IDCT,
reconstruction.  One "step" = one pass over one such clip per GPU; `value` = frames of all ranks
/ max-over-ranks device time with the clip resident in HBM; `e2e` = the same through
vcs_encode_clip_host with pinned HOST buffers (H2D of the clip and D2H of MVs, costs, flags and indices inside
the timed region; this leg does not ask for the reconstruction, so its DCT stage runs forward only).  The same numbers for the generalised true-SAD cost ride along in "sad_mode".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
N > 1: launched by torch.distributed.run, one rank per GPU (weak scaling: one clip per rank, no
data-path collective; NCCL gathers the per-shard motion vectors at the end of every step).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC_NAME = "frames/sec 1080p full-search ME+DCT/quant"
H, W, T, BS, R, GOP, QF = 1080, 1920, 60, 16, 16, 4, 50.0
STATIC_THR = 2000


def px_ops_per_p_frame(H, W, bs, R):
    """Algorithmic work of the search: one byte-difference-accumulate per byte of
    every valid candidate.  1080p/bs16/+-16: 8 590 536 candidates x 768 B = 6.598 G."""
    def n_axis(dim):
        return sum(min(p + R, dim - bs) - max(p - R, 0) + 1 for p in range(0, dim - bs + 1, bs))
    return n_axis(W) * n_axis(H) * 3 * bs * bs


def dct_bytes_per_p_frame(H, W, coef_bytes=1, recon=True):
    """Algorithmic HBM bytes of the residual/DCT/recon kernel per frame: cur 3 + ref 3 + coef
    3*coef_bytes + recon 3 per pixel (15 B/px with int16 indices, 12 B/px with int8)."""
    return H * W * (3 + 3 + 3 * coef_bytes + (3 if recon else 0))


class ClockSampler:
    """SM clock and throttle reasons sampled every ~5 ms DURING the timed region (NVML in a
    background thread; nvidia-smi -lms is too coarse for a region of tens of milliseconds)."""

    def __init__(self, index):
        self.index, self.sm, self.reasons, self.max_mhz, self._stop, self._th = index, [], set(), None, False, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nv = None
            return
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()

    def _run(self):
        nv = self._nv
        names = {getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                 getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                 getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                 getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap"}
        while not self._stop:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def stop(self):
        self._stop = True
        if self._th:
            self._th.join(timeout=2)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def bind_to_gpu_numa_node(index):
    """Pin this rank's host threads (and so its first-touched pinned buffers) to the CPUs local to its GPU:
    with 8 ranks the end-to-end path is bound by host memory / PCIe root-complex bandwidth, and a rank whose
    staging memory sits on the other socket halves its copy rate.  Best effort; returns the cpulist or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:           # NVML pads the PCI domain to 8 hex digits, sysfs uses 4
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/local_cpulist") as f:
            cpulist = f.read().strip()
        cpus = set()
        for part in cpulist.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return cpulist
    except Exception:
        pass
    return None


def make_clip(seed):
    from vcs_h264_b200 import synth
    return synth.clip(T, H, W, seed=seed)


# ------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """--impl reference: the CPU port of the reference's algorithm (oracle/vcs_oracle.c; the
    reference itself is pure Python and /root/reference does not exist on the GPU box), all host
    threads, same config/metric.  Each step = a bounded sample: the first `sample_frames` frames."""
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.build()
    sample_frames = T                                        # the whole clip: 15 I + 45 P at 1080p
    clip = make_clip(1234)[:sample_frames]
    prm = orc.symmetric_search_params(R)
    Q = orc.qtables(QF)
    cores = host_cores()          # explicit: torchrun exports OMP_NUM_THREADS=1

    def step():
        for t in range(sample_frames):
            if t % GOP:
                orc.encode_p(clip[t], clip[(t // GOP) * GOP], BS, metric=orc.METRIC_WRAP8,
                             static_thr=STATIC_THR, Q=Q, round_mode=1, simd=True, nthreads=cores, **prm)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    fps = sample_frames * args.steps / dt
    sample = (f"the whole {sample_frames}-frame clip ({sample_frames // GOP} I + "
              f"{sample_frames - sample_frames // GOP} P) per step, C port with SSE2 costs + OpenMP over macroblocks")
    emit({
        "impl": "reference", "metric": METRIC_NAME, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f64",
        "data": "synthetic", "config": workload_config(),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def workload_config():
    return {"workload": "C2: synthetic 1080p 60-frame clip (15 I + 45 P, I-P-P-P), 16x16 MB, +/-16 step-1 "
                        "full search, reference cost (wrapped u8 diff) + static test thr 2000, residual, "
                        "8x8 DCT f64, quant QF50 -> int8 indices (lossless: |idx| <= 1024/min(Q) = 102), dequant, IDCT, recon",
            "H": H, "W": W, "frames_per_clip": T, "clips_per_step": "one per GPU", "block": BS, "range": R,
            "gop": GOP, "qf": QF, "metric": "wrap8", "static_thr": STATIC_THR,
            "cache": "inputs (373 MB clip) larger than the 126 MB L2; no flush needed",
            "frames_counted": "all T frames (I-frames are stored, as in encoder.py:41-43)"}


def cpu_baseline_sample(min_seconds=10.0):
    """The CPU port on the whole 60-frame clip, repeated until >= min_seconds of CPU work."""
    from oracle import oracle as orc
    orc.build()
    clip = make_clip(1234)
    prm = orc.symmetric_search_params(R)
    Q = orc.qtables(QF)
    cores = host_cores()
    t0 = time.perf_counter()
    frames = passes = 0
    while time.perf_counter() - t0 < min_seconds:
        for t in range(T):
            if t % GOP:
                orc.encode_p(clip[t], clip[(t // GOP) * GOP], BS, metric=orc.METRIC_WRAP8,
                             static_thr=STATIC_THR, Q=Q, round_mode=1, simd=True, nthreads=cores, **prm)
            frames += 1
        passes += 1
    dt = time.perf_counter() - t0
    return {"value": frames / dt, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{passes} pass(es) over the same 60-frame clip ({frames} frames, I-frames free), "
                      f"oracle/vcs_oracle.c: SSE2 costs, OpenMP over macroblocks, {dt:.1f} s"}


def parity_spot_check(clip, first_p, strip=96):
    """Outside the timed region: the first P-frame's top `strip` rows against the CPU oracle.
    MVs / costs / quantised indices / reconstruction are integers: any mismatch is a flip."""
    from oracle import oracle as orc
    o = orc.encode_p(clip[1][:strip + R + BS], clip[0][:strip + R + BS], BS, metric=orc.METRIC_WRAP8,
                     static_thr=STATIC_THR, Q=orc.qtables(QF), round_mode=1, **orc.symmetric_search_params(R))
    n = (strip // BS) * (W // BS)
    return {"rows": strip,
            "mv_mismatch": int((first_p["mv"].astype(np.int32)[:n] != o["mv"][:n]).any(1).sum()),
            "cost_mismatch": int((first_p["cost"].view(np.uint32)[:n] != o["cost"][:n]).sum()),
            "index_flips": int((first_p["coef"][:, :strip].astype(np.float64) != o["planes"][:, :strip]).sum()),
            "recon_pixel_mismatch": int((first_p["recon"][:strip] != o["recon"][:strip]).sum()),
            "arithmetic": "float64 DCT with the reference's operation order: flips are 0 by construction"}


# ------------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import torch
    import vcs_h264_b200 as v
    from vcs_h264_b200 import _capi

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank)      # before any pinned allocation (first touch)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    clip_np = make_clip(1234 + rank)
    host_in = torch.from_numpy(clip_np).pin_memory()
    dev_in = host_in.to(dev, non_blocking=True)
    torch.cuda.synchronize()
    stream = torch.cuda.Stream(dev)          # all kernels, events and NCCL calls go on this stream
    torch.cuda.set_stream(stream)
    nP = _capi.num_p_frames(T, GOP)
    N = _capi.num_blocks(H, W, BS)

    def measure(metric):
        ce = v.ClipEncoder([H, W], block_size=BS, search="full", search_range=R, gop_len=GOP, qf=QF,
                           metric=metric, static_thr=STATIC_THR, coef_mode=v.COEF_I8_RINT, device=local_rank)
        ctx = ce.ctx
        dout = ce.alloc_device_outputs(T, want_coef=True, want_recon=True)
        mv_bytes = dout["mv"].view(torch.uint8)           # NCCL carries bytes (torch has no int16 NCCL type)
        gather = [torch.empty_like(mv_bytes) for _ in range(world)] if dist is not None else None

        def step():
            ce.encode_device(dev_in, dout, stream)
            if gather is not None:                       # per-shard results -> every rank (NCCL)
                dist.all_gather(gather, mv_bytes)

        for _ in range(args.warmup):
            step()
        barrier()
        ctx.enable_kernel_timing(True)
        l0 = ctx.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        ms = max_over_ranks(e0.elapsed_time(e1))
        launches = ctx.launch_count() - l0
        me_ms, dct_ms, ncalls = ctx.kernel_times()
        ctx.enable_kernel_timing(False)
        static_frac = float((dout["flags"] & 1).float().mean().item())

        # ---- e2e: host buffers through the C ABI, copies inside the timed region -------------
        hout = ce.alloc_host_outputs(T, want_coef=True, want_recon=False, pinned=True)
        for _ in range(max(1, args.warmup)):
            ce.encode_host(host_in, hout)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ce.encode_host(host_in, hout)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        if dist is not None:
            dist.barrier()
        same = bool(torch.equal(hout["mv"], dout["mv"].cpu()) and torch.equal(hout["coef"], dout["coef"].cpu()))
        h2d = host_in.numel()
        d2h = sum(hout[k].numel() * hout[k].element_size() for k in ("mv", "cost", "flags", "coef"))
        first_p = {k: dout[k][0].cpu().numpy() for k in ("mv", "cost", "coef", "recon")}
        return dict(first_p=first_p, ms=ms, launches=launches, me_ms=me_ms / max(ncalls, 1), dct_ms=dct_ms / max(ncalls, 1),
                    e2e_fps=world * T * args.steps / dt, e2e_ms=1e3 * dt / args.steps, h2d=h2d, d2h=d2h,
                    clocks=clocks, static_frac=static_frac, host_equals_device=same)

    ctx0 = v.runtime.get_context(local_rank)
    ctx0.set_stream(stream.cuda_stream)
    mb = {}
    if rank == 0:
        for which, name in ((0, "vabsdiff4_acc"), (1, "iadd3"), (2, "lop3"), (3, "imad"), (4, "idp4a"),
                            (5, "wrap8_triple_words"), (6, "vabsdiff4_with_lds")):
            rate, mhz = ctx0.microbench(which, 4000)
            mb[name] = {"warp_instr_per_s": rate, "sm_mhz": mhz,
                        "per_clk_per_sm": rate / (mhz * 1e6) / ctx0.device_info()["sm_count"]}
    wrap = measure(_capi.METRIC_WRAP8)
    sad = measure(_capi.METRIC_SAD) if not args.skip_sad else None

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    hbm_peak, hbm_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback")
    pxops = px_ops_per_p_frame(H, W, BS, R) * nP           # per ME launch (one clip)
    peak_pxops = mb["vabsdiff4_acc"]["warp_instr_per_s"] * 32 * 4

    # DRAM bytes per launch from the ncu --set full capture of this command (profiles/r1_ncu_full_summary.csv)
    ME_TRAFFIC, DCT_TRAFFIC = 381.6e6, 1033.6e6

    sm_count = ctx0.device_info()["sm_count"]

    def me_roof(m, instr_per_word=1):
        ach = pxops / (m["me_ms"] * 1e-3)
        # issue-slot view: the cost needs `instr_per_word` warp instructions per 4 bytes at best; one
        # scheduler issues one instruction per clock (4 per SM)
        issue_peak = sm_count * 4 * (wrap["clocks"]["sm_mhz"] or 1965.0) * 1e6
        issue_frac = (pxops / 4 / 32 * instr_per_word) / (m["me_ms"] * 1e-3) / issue_peak
        return {"min_instr_per_word": instr_per_word, "frac_of_issue_slots": issue_frac,"bound": "int32", "kernel": "me_tiled_kernel (one launch = 45 P-frames)", "achieved": ach / 1e9,
                "peak": peak_pxops / 1e9, "unit": "Gpxop/s", "frac": ach / peak_pxops, "traffic": ME_TRAFFIC,
                "traffic_unit": "bytes of DRAM traffic per launch (ncu); algorithmic input 373.2e6",
                "ms_per_launch": m["me_ms"],
                "peak_source": "VABSDIFF4.U8.ACC issue rate measured in this run x 32 lanes x 4 bytes"}

    def dct_roof(m):
        b = dct_bytes_per_p_frame(H, W) * nP
        ach = b / (m["dct_ms"] * 1e-3) / 1e9
        return {"bound": "hbm", "kernel": "dct_stage_kernel (one launch = 45 P-frames)", "achieved": ach,
                "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": DCT_TRAFFIC,
                "algorithmic_bytes": b, "note": "FP64-pipe bound before HBM (96 DFMA/px, bit-exact float64 DCT)",
                "ms_per_launch": m["dct_ms"], "peak_source": hbm_src}

    line = {
        "metric": METRIC_NAME, "value": world * T * args.steps / (wrap["ms"] * 1e-3), "unit": "frames/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": wrap["ms"] / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f64", "data": "synthetic",
        "config": workload_config(), "clocks": wrap["clocks"],
        "e2e": {"value": wrap["e2e_fps"], "unit": "frames/s", "h2d_bytes_per_step": wrap["h2d"],
                "d2h_bytes_per_step": wrap["d2h"], "ms_per_step": wrap["e2e_ms"],
                "host_equals_device": wrap["host_equals_device"]},
        "gpu_launches": wrap["launches"],
        "roofline": me_roof(wrap, 3), "roofline_dct": dct_roof(wrap),
        "static_fraction": wrap["static_frac"], "microbench": mb, "host_numa_cpus_rank0": numa,
    }
    if sad is not None:
        line["sad_mode"] = {"value": world * T * args.steps / (sad["ms"] * 1e-3), "unit": "frames/s",
                            "ms_per_step": sad["ms"] / args.steps, "e2e": sad["e2e_fps"],
                            "roofline": me_roof(sad), "clocks": sad["clocks"]}
    if world == 1 and not args.skip_cpu:
        line["cpu_baseline"] = cpu_baseline_sample()
        line["parity_check"] = parity_spot_check(clip_np, wrap["first_p"])
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


_JSON_OUT = None


def emit(line):
    """The one JSON line goes to the process's ORIGINAL stdout; fd 1 itself is pointed at stderr for the whole run
    (main()) so that library chatter -- NCCL's version banner, OpenMP notices -- can never land in front of it."""
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--skip-sad", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
