#!/usr/bin/env python
"""bench.py - headline benchmark: 4K 10-bit VVC reconstruction + in-loop filtering, Mpix/s per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json config 5): every picture of a ring of distinct synthetic
3840x2160 10-bit 4:2:0 pictures goes through the whole hot path in the reference's stage order
    INTER (MC, bi-pred, DMVR, BDOF, PROF, GPM) -> residual (LFNST + inverse transforms + add_residual)
    -> inverse LMCS -> deblock V -> deblock H -> SAO -> ALF / CC-ALF.
One "step" = one pass over the ring.  The ring (pictures, reference pictures, coefficients) is several
times larger than the 126 MB L2.

Ours: every stage is a CUDA kernel of libvvcdsp_cuda.so called through the C ABI.  `value` has pictures
and descriptors resident in HBM; `e2e` goes through vvc_cuda_recon_frame_host with pinned HOST buffers
(H2D of references, records and coefficients + kernels + D2H of the pictures inside the timed region).
Reference arm (--impl reference) and the cpu_baseline leg: the reference's own C table entries
(oracle/_ref/libvvcref.so, else the oracle port) on the host cores, one picture per thread.
Multi-GPU: independent streams, one process per GPU, no data-path collective.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from ffvvc_b200 import abi, streams, synth  # noqa: E402

METRIC = "vvc_4k10_recon_loopfilter_mpix_per_s"
UNIT = "Mpix/s"
STAGES = ["inter", "residual", "lmcs", "deblock_v", "deblock_h", "sao", "alf"]
# algorithmic bytes per luma pixel moved by ONE launch of each stage (SURVEY.md 8(d), 4:2:0 10 bit):
# inter all-bi 9; residual 8 B/sample x 1.5; LMCS luma r+w; each filter pass reads + writes 3 planes
ALGO_BYTES_PER_LUMA_PX = {"inter": 9.0, "residual": 12.0, "lmcs": 4.0, "deblock_v": 6.0, "deblock_h": 6.0, "sao": 6.0, "alf": 6.0}
CHAIN_ALGO_BYTES_PER_LUMA_PX = 43.0     # 9 + 12 + 4 + 18 (deblock V+H counted once), SURVEY.md config 5


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------
# Synthetic workload
# ---------------------------------------------------------------------------------------------
class Inputs:
    """Host-side inputs for `distinct` different pictures; ring slot k reuses content k % distinct at
    its own addresses (so the ring still defeats L2)."""

    def __init__(self, width, height, seed=12345, distinct=2, lfnst_set_of=None):
        self.w, self.h, self.distinct = width, height, distinct
        self.g1 = abi.FrameGeom(width, height)
        gd = abi.FrameGeom(width, height, batch=distinct)
        self.ref_planes = synth.struct_planes(gd, seed=seed)          # reference pictures (distinct contents)
        self.pbs, self.tbs, self.coeffs, self.maps, self.sao, self.alf = [], [], [], [], [], []
        self.quant, self.win_tbs, self.win, self.scaling = [], [], [], None
        profs = []
        for i in range(distinct):
            pbs, self.wp, prof = synth.pb_list(self.g1, n_refs=2, seed=seed + 10 * i + 1)
            pbs["prof"] += sum(len(q) for q in profs)          # every content indexes its own part of one PROF table
            profs.append(prof)
            self.pbs.append(pbs)
            tbs, co = synth.tb_list(self.g1, seed=seed + 10 * i + 2, lfnst_set_of=lfnst_set_of, extras=False, saturate=False)
            # the residual stage starts from quantised levels (TransCoeffLevel) like ff_vvc_reconstruct does: dequant()
            # is part of the stage in every arm.  Dense int32 (the reference's tb->coeffs layout) for the reference arm
            # and the device-resident number, the 16-bit window layout for the upload of the end-to-end number.
            tbs = synth.tb_for_window(tbs)
            co = (co >> 4).astype(np.int32)
            q, sl = synth.tb_quant(tbs, seed=seed + 10 * i + 6, scaling=True, qp_lo=22, qp_hi=42)
            wt, win = abi.pack_window16(tbs, co)
            self.tbs.append(tbs)
            self.coeffs.append(co)
            self.quant.append(q)
            self.scaling = sl if i == 0 else self.scaling
            self.win_tbs.append(wt)
            self.win.append(win)
            self.maps.append(synth.deblock_maps(self.g1, seed=seed + 10 * i + 3, qp_base=27, qp_span=16))
            self.sao.append(synth.sao_params(self.g1, seed=seed + 10 * i + 4))
            alf, self.sets = synth.alf_params(self.g1, seed=seed + 10 * i + 5)
            self.alf.append(alf)
        self.prof = np.concatenate(profs)
        _, self.inv_lut = synth.lmcs_luts(10, seed=seed + 7)

    def records(self, k, n_ref_slots, pic):
        """Prediction records of ring picture k: reference slots (2k, 2k+1) mod ring, destination `pic`."""
        pbs = self.pbs[k % self.distinct].copy()
        pbs["ref"] = (2 * k + pbs["ref"]) % n_ref_slots
        pbs["pic"] = pic
        return pbs

    def frame_bytes(self):
        return sum(self.g1.plane_wh(c)[0] * self.g1.plane_wh(c)[1] * 2 for c in range(3))


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def result(self):
        self.stop_flag = True
        if not self.samples and self.nv:           # a timed region shorter than one sampling period: sample right at its end
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            except Exception:
                pass
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's C (or the oracle port) on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_lib():
    FP, MP = C.POINTER(abi.VVCCudaFrame), C.POINTER(abi.VVCCudaDeblockMaps)
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libvvcref.so")
    if os.path.exists(ref_so):
        lib, kind, pre = C.CDLL(ref_so), "reference", "vvcref_"
    else:
        lib, kind, pre = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so")), "port", "vvco_"
    fns = {}
    sig = {"inter_frame": [FP, FP, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p],
           "itx_frame_q": [FP, C.POINTER(abi.VVCCudaCoeffs), C.c_void_p, C.c_int, C.c_int],
           "lmcs_frame": [FP, C.c_void_p, C.c_void_p],
           "deblock_frame": [FP, FP, MP, C.c_int], "sao_frame": [FP, FP, C.c_void_p],
           "alf_frame": [FP, FP, C.c_void_p, C.c_void_p, C.c_int]}
    for name, args in sig.items():
        f = getattr(lib, pre + name)
        f.argtypes, f.restype = args, None
        fns[name] = f
    return kind, fns


def cpu_reconstruct(fns, inp, i, refs, scratch):
    """One picture (content i) through the reference C stage by stage; returns the output planes."""
    g, gr = inp.g1, abi.FrameGeom(inp.w, inp.h, batch=inp.distinct)
    cur, a, b = scratch
    pbs = inp.pbs[i]
    fns["inter_frame"](abi.frame_from_numpy(g, cur), abi.frame_from_numpy(gr, refs), pbs.ctypes.data, len(pbs),
                       inp.wp.ctypes.data, inp.prof.ctypes.data, None)
    co = abi.coeffs_desc(inp.coeffs[i].ctypes.data, inp.coeffs[i].size, abi.COEFF_DENSE32, inp.quant[i].ctypes.data, inp.scaling.ctypes.data)
    fns["itx_frame_q"](abi.frame_from_numpy(g, cur), C.byref(co), inp.tbs[i].ctypes.data, len(inp.tbs[i]), 15)
    fns["lmcs_frame"](abi.frame_from_numpy(g, cur), inp.inv_lut.ctypes.data, None)
    md = abi.deblock_maps_desc(g, inp.maps[i])
    fns["deblock_frame"](abi.frame_from_numpy(g, a), abi.frame_from_numpy(g, cur), C.byref(md), 1)
    fns["deblock_frame"](abi.frame_from_numpy(g, b), abi.frame_from_numpy(g, a), C.byref(md), 0)
    fns["sao_frame"](abi.frame_from_numpy(g, a), abi.frame_from_numpy(g, b), inp.sao[i].ctypes.data)
    fns["alf_frame"](abi.frame_from_numpy(g, b), abi.frame_from_numpy(g, a), inp.alf[i].ctypes.data, inp.sets.ctypes.data, 0)
    return b


def run_cpu(inp, steps, warmup, threads):
    """Each step: `threads` host threads each reconstruct one 4K picture with the reference C."""
    kind, fns = cpu_lib()
    scratch = [[abi.alloc_planes(inp.g1) for _ in range(3)] for _ in range(threads)]

    errors = []

    def work(t):
        try:
            cpu_reconstruct(fns, inp, t % inp.distinct, inp.ref_planes, scratch[t])
        except Exception as e:          # a silent thread death would fake a fast CPU
            errors.append(e)

    def one_step():
        ts = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        if errors:
            raise errors[0]
        return time.perf_counter() - t0

    for _ in range(warmup):
        one_step()
    total = sum(one_step() for _ in range(steps))
    return kind, inp.w * inp.h * threads * steps / total / 1e6, total / steps * 1e3


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--frames", type=int, default=16, help="pictures in the ring = pictures per step")
    ap.add_argument("--group", type=int, default=16, help="pictures per launch (independent streams batched into one launch per stage)")
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--seed", type=int, default=12345, help="seed of the synthetic inputs (rank r uses seed + r)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity gate against the reference C")
    ap.add_argument("--inter-tma", type=int, default=-1, help="VVC_CUDA_OPT_INTER_TMA (default: the library's)")
    ap.add_argument("--ref-pad", type=int, default=128, help="VVC_CUDA_OPT_REF_PAD: margin of replicated samples around the device-resident DPB ring")
    ap.add_argument("--quick", action="store_true", help="skip the one-picture-per-launch and the >1 s runs")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    threads = args.cpu_threads or (os.cpu_count() or 1)
    workload = ("recon_4k: INTER (MC/bi/DMVR/BDOF/PROF/GPM) -> residual (dequant+LFNST+itx+add) -> inverse LMCS -> deblock V+H -> SAO -> "
                "ALF/CC-ALF on %dx%d 10-bit 4:2:0, 100%% inter area, coded fraction 1.0") % (args.width, args.height)

    lfnst_set_of = synth.lfnst_set_of              # LFNST set of an intra mode, from the generated tables
    # the same `config` in both arms: what is measured; how each arm runs it goes under "run"
    config = {"workload": workload, "width": args.width, "height": args.height, "bit_depth": 10, "chroma_format": "4:2:0",
              "coded_fraction": 1.0, "stages": STAGES,
              "l2": "inputs larger than L2: every step streams a ring of pictures, references and coefficients several times the 126 MB L2, no flush needed",
              "parallelism": "independent streams x%d, no collective" % args.gpus}

    if args.impl == "reference":
        if rank != 0:
            return 0
        inp = Inputs(args.width, args.height, seed=args.seed, distinct=2, lfnst_set_of=lfnst_set_of)
        kind, mpix, ms = run_cpu(inp, args.steps, max(args.warmup, 1), threads)
        _, mpix1, ms1 = run_cpu(inp, 1, 0, 1)
        sample = "%d pictures per step (one %dx%d picture per host thread) of the same synthetic workload" % (threads, args.width, args.height)
        line = {
            "impl": "reference", "metric": METRIC, "value": mpix, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u16", "data": "synthetic", "config": config,
            "run": {"cpu_threads": threads, "pictures_per_step": threads},
            "cpu_baseline": {"value": mpix, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample,
                             "one_thread": {"value": mpix1, "unit": UNIT, "ms_per_picture": ms1}, "cpu_model": cpu_model(),
                             "simd": "C only: the reference's x86 assembly needs nasm, which this image lacks, so its AVX2 MC/SAO/ALF is not represented"},
            "e2e": {"value": mpix, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line))
        return 0

    import torch
    from ffvvc_b200 import device, lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = "cuda:%d" % local_rank
    if world > 1:
        import torch.distributed as dist
        # NCCL may print its version banner on stdout: keep stdout for the one JSON line
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device(dev))
            dist.barrier()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    frames, group = args.frames, max(1, min(args.group, args.frames))
    while frames % group:
        group -= 1
    inp = Inputs(args.width, args.height, seed=streams.seed_of_rank(args.seed, rank), distinct=2, lfnst_set_of=lfnst_set_of)
    g1 = inp.g1
    gring = abi.FrameGeom(args.width, args.height, batch=frames)
    ggrp = abi.FrameGeom(args.width, args.height, batch=group)
    ctx = lib.Context(local_rank)
    if args.inter_tma >= 0:
        ctx.set_option(abi.OPT_INTER_TMA, args.inter_tma)
    stream = ctx.torch_stream()
    torch.cuda.set_stream(stream)      # uploads, events and kernels all on the context's stream

    # ---- device-resident rings --------------------------------------------------------------------
    reps = frames // inp.distinct + 1
    ring_planes = [np.ascontiguousarray(np.concatenate([p] * reps)[:frames]) for p in inp.ref_planes]
    refs = device.DeviceFrames(gring, device=dev, planes=ring_planes, pad=args.ref_pad)       # DPB ring (frames pictures > L2)
    if args.ref_pad:
        ctx.pad_frame(refs.desc, args.ref_pad)
        ctx.set_option(abi.OPT_REF_PAD, args.ref_pad)
    cur = device.DeviceFrames(gring, device=dev)
    out = device.DeviceFrames(gring, device=dev)
    tmp_a = device.DeviceFrames(ggrp, device=dev)
    tmp_b = device.DeviceFrames(ggrp, device=dev)
    keep = []

    def up(a):
        t, p = device.to_device(a, dev)
        keep.append(t)
        return p

    p_wp, p_prof, p_sets, p_lut, p_sl = up(inp.wp), up(inp.prof), up(inp.sets), up(inp.inv_lut), up(inp.scaling)
    n_ctb = g1.ctb_count

    def sub_frame(df, k0, n):
        f = abi.VVCCudaFrame()
        C.memmove(C.byref(f), C.byref(df.desc), C.sizeof(f))
        for c in range(3):
            f.data[c] = df.desc.data[c] + k0 * df.desc.batch_stride[c]
        f.batch = n
        return f

    def build_groups(group):
        """Descriptors of the ring, `group` pictures (independent streams) per launch."""
        groups = []
        for k0 in range(0, frames, group):
            ks = list(range(k0, k0 + group))
            pbs = np.concatenate([inp.records(k, frames, j) for j, k in enumerate(ks)])
            tb_parts, co_parts, off = [], [], 0
            for j, k in enumerate(ks):
                t = inp.tbs[k % inp.distinct].copy()
                t["pic"] = j
                t["coeff_offset"] += off
                off += len(inp.coeffs[k % inp.distinct])
                tb_parts.append(t)
                co_parts.append(inp.coeffs[k % inp.distinct])
            tbs, coeffs = np.concatenate(tb_parts), np.concatenate(co_parts)
            quant = np.concatenate([inp.quant[k % inp.distinct] for k in ks])
            md = abi.VVCCudaDeblockMaps()
            for d in range(2):
                for c in range(3):
                    rows, pitch = abi.deblock_map_shape(g1, d, c)
                    arr = np.concatenate([inp.maps[k % inp.distinct][d][c] for k in ks])
                    md.edge[d][c] = up(arr)
                    md.pitch[d][c], md.rows[d][c], md.size[d][c] = pitch, rows, rows * pitch
            groups.append(dict(
                cur=sub_frame(cur, k0, group), out=sub_frame(out, k0, group), ta=sub_frame(tmp_a, 0, group), tb=sub_frame(tmp_b, 0, group),
                pbs=up(pbs), n_pbs=len(pbs), tbs=up(tbs), n_tbs=len(tbs),
                coeffs=abi.coeffs_desc(up(coeffs), len(coeffs), abi.COEFF_DENSE32, up(quant), p_sl), md=md,
                sao=up(np.concatenate([inp.sao[k % inp.distinct] for k in ks])),
                alf=up(np.concatenate([inp.alf[k % inp.distinct] for k in ks]))))
        return groups

    groups = build_groups(group)
    # kernels per picture group: inter = 8 (classify, four thread-per-patch class kernels, DMVR search, two warp-per-record kernels),
    # residual = 4 (size binning, thread-per-block kernel for 2x2..4x4, warp-per-TB kernel, generic kernel over the
    # blocks those leave), every other stage 1
    launches_per_group = len(STAGES) + 10
    launches_per_step = len(groups) * launches_per_group          # re-counted from the library's own counter after the warm-up

    def step(events=None, groups=groups):
        for gi, g in enumerate(groups):
            ev = events[gi] if events is not None else None
            if ev: ev[0].record()
            ctx.inter_frame(g["cur"], refs.desc, g["pbs"], g["n_pbs"], p_wp, p_prof, None)
            if ev: ev[1].record()
            ctx.itx_frame_q(g["cur"], g["coeffs"], g["tbs"], g["n_tbs"], 15)
            if ev: ev[2].record()
            ctx.lmcs_frame(g["cur"], p_lut, None)
            if ev: ev[3].record()
            ctx.deblock_frame(g["ta"], g["cur"], g["md"], 1)
            if ev: ev[4].record()
            ctx.deblock_frame(g["tb"], g["ta"], g["md"], 0)
            if ev: ev[5].record()
            ctx.sao_frame(g["ta"], g["tb"], g["sao"])
            if ev: ev[6].record()
            ctx.alf_frame(g["out"], g["ta"], g["alf"], p_sets, 0)
            if ev: ev[7].record()

    def timed(n_steps, groups):
        """n_steps passes over the ring, CUDA events on the context's stream; max over ranks."""
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a.record()
        for _ in range(n_steps):
            step(None, groups)
        b.record()
        barrier()
        ms = a.elapsed_time(b)
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            torch.cuda.synchronize()

    # Warm-up: at least W (>= 3) steps, and at least 0.5 s of them.  After tens of seconds of host-side input synthesis the
    # GPU has dropped to an idle power state; the first tens of milliseconds after it wakes were measured slower in the
    # latency-sensitive residual stage (1.37 instead of 1.10 ms per launch) while a later 1.5 s run of the same process
    # gave the steady figure, so the untimed warm-up runs until the clocks have settled.
    warm_steps, t_warm = 0, time.perf_counter()
    while warm_steps < max(args.warmup, 3) or time.perf_counter() - t_warm < 0.5:
        step()
        warm_steps += 1
        if warm_steps % 8 == 0:
            torch.cuda.synchronize()
    l_w = ctx.launches
    step()
    launches_per_step = ctx.launches - l_w                        # kernels this library launched for one step
    barrier()

    # ---- timed region: K steps, device events; per-kernel events ride along --------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ctx.launches
    nst = len(STAGES)
    evs = [[[torch.cuda.Event(enable_timing=True) for _ in range(nst + 1)] for _ in groups] for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_start.record()
    for s in range(args.steps):
        step(evs[s])
    t_end.record()
    barrier()
    clocks = sampler.result()
    elapsed_ms = t_start.elapsed_time(t_end)
    gpu_launches = ctx.launches - l0
    ctx.sync()
    assert gpu_launches == launches_per_step * args.steps, (gpu_launches, launches_per_step)

    stage_ms = np.zeros(nst)
    for s in range(args.steps):
        for g in evs[s]:
            for i in range(nst):
                stage_ms[i] += g[i].elapsed_time(g[i + 1])
    stage_ms /= args.steps * len(groups)          # average duration of one launch of each stage

    # every rank runs its own streams (no data-path collective): units add up, the slowest rank's device time counts
    luma_px_per_step = args.width * args.height * frames
    dist_mod = None
    if world > 1:
        import torch.distributed as dist_mod
    total_px, elapsed_ms = streams.aggregate(luma_px_per_step * args.steps, elapsed_ms, dist_mod, dev if world > 1 else None)
    value = streams.throughput_mpix(total_px, elapsed_ms)

    # ---- roofline of the dominant kernel ------------------------------------------------------------
    peak, peak_src = load_peaks()
    dom = int(np.argmax(stage_ms))
    px_per_launch = args.width * args.height * group
    algo_bytes = ALGO_BYTES_PER_LUMA_PX[STAGES[dom]] * px_per_launch
    achieved = algo_bytes / (stage_ms[dom] * 1e-3) / 1e9
    per_stage = {n: {"ms_per_launch": float(v), "achieved_gbs": ALGO_BYTES_PER_LUMA_PX[n] * px_per_launch / (v * 1e-3) / 1e9,
                     "frac": ALGO_BYTES_PER_LUMA_PX[n] * px_per_launch / (v * 1e-3) / 1e9 / peak}
                 for n, v in zip(STAGES, stage_ms)}
    chain_gbs = CHAIN_ALGO_BYTES_PER_LUMA_PX * luma_px_per_step * args.steps / (elapsed_ms * 1e-3) / 1e9 * (1.0 if world == 1 else 1.0)
    # measured DRAM traffic of the dominant stage, per launch like `achieved`: dram__bytes_read + dram__bytes_write of its kernels in
    # one `ncu --set full` capture at the bench's 16 pictures per launch (profiles/r02_traffic.json, from r02_recon16_ncu_full.csv)
    traffic, traffic_src = None, None
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", name)))["per_stage"][STAGES[dom]]
            traffic = (tr["dram_read_mb_per_picture"] + tr["dram_write_mb_per_picture"]) * 1e6 * group * (args.width * args.height) / (3840 * 2160)
            traffic_src = "profiles/" + name
            break
        except Exception:
            pass
    roofline = {
        "bound": "hbm", "kernel": STAGES[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": algo_bytes, "stages": per_stage,
        "chain": {"algorithmic_bytes_per_luma_px": CHAIN_ALGO_BYTES_PER_LUMA_PX, "achieved_gbs": chain_gbs, "frac": chain_gbs / peak},
    }

    # ---- the same ring with ONE picture per launch (BASELINE config 5 as worded: one stream per GPU) and a run of
    # ---- more than a second (the headline region is a burst of ~0.1 s) ----------------------------------
    extra = {}
    if not args.quick:
        if group > 1:
            g_one = build_groups(1)
            step(None, g_one)
            ms1 = timed(min(args.steps, 5), g_one)
            extra["one_picture_per_launch"] = {"value": luma_px_per_step * min(args.steps, 5) * world / (ms1 * 1e-3) / 1e6, "unit": UNIT,
                                               "ms_per_step": ms1 / min(args.steps, 5)}
            del g_one
        n_long = max(args.steps, int(np.ceil(1500.0 / (elapsed_ms / args.steps))))
        sampler2 = ClockSampler(local_rank)
        sampler2.start()
        ms_long = timed(n_long, groups)
        extra["sustained"] = {"value": luma_px_per_step * n_long * world / (ms_long * 1e-3) / 1e6, "unit": UNIT, "steps": n_long,
                              "seconds": ms_long * 1e-3, "clocks": sampler2.result()}

    # ---- e2e: pinned host buffers through vvc_cuda_recon_frame_host -----------------------------------
    e2e = None
    if not args.no_e2e:
        pin_keep = []

        def pin(a):
            t = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).pin_memory()
            pin_keep.append(t)
            return t.data_ptr(), t.numel()

        h_refs = [torch.from_numpy(p.view(np.int16)).pin_memory() for p in ring_planes]
        h_out = [torch.empty_like(t).pin_memory() for t in h_refs]
        f_refs = abi.frame_desc(gring, [t.data_ptr() for t in h_refs], [t.stride(1) * 2 for t in h_refs], [t.stride(0) * 2 for t in h_refs])
        f_out = abi.frame_desc(gring, [t.data_ptr() for t in h_out], [t.stride(1) * 2 for t in h_out], [t.stride(0) * 2 for t in h_out])
        (hp_wp, b_wp), (hp_prof, b_prof), (hp_sets, b_sets), (hp_lut, b_lut) = pin(inp.wp), pin(inp.prof), pin(inp.sets), pin(inp.inv_lut)
        hp_sl, b_sl = pin(inp.scaling)
        per_content = []
        for i in range(inp.distinct):
            hmd = abi.VVCCudaDeblockMaps()
            nb = 0
            for d in range(2):
                for c in range(3):
                    rows, pitch = abi.deblock_map_shape(g1, d, c)
                    hmd.edge[d][c], n = pin(inp.maps[i][d][c])
                    nb += n
                    hmd.pitch[d][c], hmd.rows[d][c], hmd.size[d][c] = pitch, rows, rows * pitch
            (p_co, b_co), (p_tb, b_tb), (p_sa, b_sa), (p_al, b_al) = pin(inp.coeffs[i]), pin(inp.tbs[i]), pin(inp.sao[i]), pin(inp.alf[i])
            (p_win, b_win), (p_wtb, b_wtb), (p_q, b_q) = pin(inp.win[i]), pin(inp.win_tbs[i]), pin(inp.quant[i])
            per_content.append(dict(md=hmd, co=p_co, n_co=len(inp.coeffs[i]), tb=p_tb, n_tb=len(inp.tbs[i]), sao=p_sa, alf=p_al,
                                    win=p_win, n_win=len(inp.win[i]), wtb=p_wtb, quant=p_q,
                                    bytes=nb + b_sa + b_al + b_q + b_sl, bytes_dense=b_co + b_tb, bytes_win=b_win + b_wtb))
        descs = (abi.VVCCudaReconDesc * frames)()            # 16-bit window layout (the upload format of the product)
        descs_dense = (abi.VVCCudaReconDesc * frames)()      # the reference's dense int32 layout, for comparison
        h2d_dense_extra = 0
        h2d = sum(t.numel() * 2 for t in h_refs)
        d2h = sum(t.numel() * 2 for t in h_out)
        for k in range(frames):
            pc = per_content[k % inp.distinct]
            p_pb, b_pb = pin(inp.records(k, frames, 0))
            d = descs[k]
            d.pbs, d.n_pbs, d.wp, d.n_wp, d.prof, d.n_prof = p_pb, len(inp.pbs[k % inp.distinct]), hp_wp, len(inp.wp), hp_prof, len(inp.prof)
            d.log2_transform_range = 15
            d.coeffs, d.n_coeffs, d.tbs, d.n_tbs = pc["win"], pc["n_win"], pc["wtb"], pc["n_tb"]
            d.coeff_format, d.quant, d.scaling = abi.COEFF_WINDOW16, pc["quant"], hp_sl
            d.ref_slots = (1 << ((2 * k) % frames)) | (1 << ((2 * k + 1) % frames))     # what Inputs.records() references
            d.lmcs_inv_lut = hp_lut
            d.inloop.deblock = C.pointer(pc["md"])
            d.inloop.sao, d.inloop.alf, d.inloop.alf_sets = pc["sao"], pc["alf"], hp_sets
            h2d += b_pb + b_wp + b_prof + b_lut + b_sets + pc["bytes"] + pc["bytes_win"]
            h2d_dense_extra += pc["bytes_dense"] - pc["bytes_win"]
            C.memmove(C.byref(descs_dense[k]), C.byref(d), C.sizeof(d))
            dd = descs_dense[k]
            dd.coeffs, dd.n_coeffs, dd.tbs, dd.coeff_format = pc["co"], pc["n_co"], pc["tb"], abi.COEFF_DENSE32
        # the product's upload format: ONE pinned arena per picture (vvc_cuda_recon_arena_bind), so a picture's descriptors,
        # records and levels go up as a single copy; `descs` above (one pinned array per table) stays as the comparison
        descs_sep = descs
        descs = (abi.VVCCudaReconDesc * frames)()
        arena_keep = []

        def arena_alloc(n):
            t = torch.empty(n + 256, dtype=torch.uint8).pin_memory()
            arena_keep.append(t)
            return t, (t.data_ptr() + 255) & ~255

        arena_bytes = 0
        for k in range(frames):
            i = k % inp.distinct
            d, m, _ = abi.recon_arena(lib.load(), g1, arena_alloc, pbs=inp.records(k, frames, 0), wp=inp.wp, prof=inp.prof, tbs=inp.win_tbs[i],
                                      coeffs=inp.win[i], coeff_format=abi.COEFF_WINDOW16, quant=inp.quant[i], scaling=inp.scaling,
                                      inv_lut=inp.inv_lut, maps=[[inp.maps[i][dr][c][0] for c in range(3)] for dr in range(2)],
                                      sao=inp.sao[i], alf=inp.alf[i], sets=inp.sets,
                                      ref_slots=(1 << ((2 * k) % frames)) | (1 << ((2 * k + 1) % frames)))
            arena_keep.append(m)
            C.memmove(C.byref(descs[k]), C.byref(d), C.sizeof(d))
            descs[k].inloop.deblock = C.pointer(m)
            arena_bytes += int(d.arena_bytes)
        h2d = sum(t.numel() * 2 for t in h_refs) + arena_bytes
        e_steps = max(2, min(args.steps, 8))
        ctx.recon_frame_host(f_out, f_refs, descs)

        def timed_host_steps(call, sync_each):
            barrier()
            es, ee = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            es.record()
            for _ in range(e_steps):
                call(f_out, f_refs, descs)
            if not sync_each:
                ctx.sync()                                  # the context stream and both copy streams: every output picture is in host memory
            ee.record()
            barrier()
            wall_ms = (time.perf_counter() - t0) * 1e3
            ms = max(es.elapsed_time(ee), wall_ms)          # the host waited for the last copy: wall clock is the honest bound
            if world > 1:
                import torch.distributed as dist
                t = torch.tensor([ms], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            return ms
        # a step = one call of 16 pictures; the calls are queued back to back (vvc_cuda_recon_frame_host_async) and the host
        # waits once, after the last step's last copy-out - the way a decoder that delivers pictures continuously drives the entry
        e_ms = timed_host_steps(ctx.recon_frame_host_async, False)
        s_ms = timed_host_steps(ctx.recon_frame_host, True)       # every call waits for its own last copy-out: the pipeline fills and drains per step
        e2e = {"value": luma_px_per_step * e_steps * world / (e_ms * 1e-3) / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e_steps,
               "h2d_gbs_per_rank": h2d * e_steps / (e_ms * 1e-3) / 1e9, "d2h_gbs_per_rank": d2h * e_steps / (e_ms * 1e-3) / 1e9,
               "api": "vvc_cuda_recon_frame_host_async, one call per step, vvc_cuda_sync after the last step (pinned host reference pictures; one pinned arena per picture holding its records, quantised levels in the 16-bit window layout and filter parameters; output pictures copied back)",
               "synchronous_calls": {"value": luma_px_per_step * e_steps * world / (s_ms * 1e-3) / 1e6, "unit": UNIT,
                                     "what": "vvc_cuda_recon_frame_host: every step waits for its own last copy-out"}}
        # the same call with the DPB already in HBM (reference pictures are earlier outputs in a decoder): only the
        # per-picture records / coefficients / filter parameters go up, the output pictures come back
        ctx.recon_frame_host(f_out, refs.desc, descs)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            ctx.recon_frame_host_async(f_out, refs.desc, descs)
        ctx.sync()
        barrier()
        r_ms = (time.perf_counter() - t0) * 1e3
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([r_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            r_ms = float(t.item())
        e2e["dpb_resident"] = {"value": luma_px_per_step * e_steps * world / (r_ms * 1e-3) / 1e6, "unit": UNIT,
                               "h2d_bytes_per_step": int(h2d - sum(t.numel() * 2 for t in h_refs)), "d2h_bytes_per_step": int(d2h)}
        # ... and with the output ring staying in HBM as well (a decoder's output pictures ARE its later references; what
        # leaves the device is what the display or the encoder asks for): only the per-picture arenas move
        ctx.recon_frame_host(out.desc, refs.desc, descs)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            ctx.recon_frame_host_async(out.desc, refs.desc, descs)
        ctx.sync()
        barrier()
        o_ms = (time.perf_counter() - t0) * 1e3
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([o_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            o_ms = float(t.item())
        e2e["dpb_and_output_resident"] = {"value": luma_px_per_step * e_steps * world / (o_ms * 1e-3) / 1e6, "unit": UNIT,
                                          "h2d_bytes_per_step": int(arena_bytes), "d2h_bytes_per_step": 0,
                                          "h2d_gbs_per_rank": arena_bytes * e_steps / (o_ms * 1e-3) / 1e9}
        # one pinned array per table instead of one arena per picture (about 20 copies per picture)
        ctx.recon_frame_host(f_out, f_refs, descs_sep)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            ctx.recon_frame_host(f_out, f_refs, descs_sep)
        barrier()
        s_ms = (time.perf_counter() - t0) * 1e3
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([s_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            s_ms = float(t.item())
        e2e["one_copy_per_table"] = {"value": luma_px_per_step * e_steps * world / (s_ms * 1e-3) / 1e6, "unit": UNIT}
        # the PCIe / host-memory ceiling of this box for the same bytes: the copies alone, both directions at once, no kernels
        hp_in = torch.empty(int(h2d) // frames, dtype=torch.uint8).pin_memory()
        hp_out = torch.empty(int(d2h) // frames, dtype=torch.uint8).pin_memory()
        dv_in, dv_out = torch.empty_like(hp_in, device=dev), torch.empty_like(hp_out, device=dev)
        s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps * frames):
            with torch.cuda.stream(s_in):
                dv_in.copy_(hp_in, non_blocking=True)
            with torch.cuda.stream(s_out):
                hp_out.copy_(dv_out, non_blocking=True)
        barrier()
        c_ms = (time.perf_counter() - t0) * 1e3
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([c_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            c_ms = float(t.item())
        e2e["copies_only_ceiling"] = {"value": luma_px_per_step * e_steps * world / (c_ms * 1e-3) / 1e6, "unit": UNIT,
                                      "h2d_gbs_per_rank": h2d * e_steps / (c_ms * 1e-3) / 1e9, "d2h_gbs_per_rank": d2h * e_steps / (c_ms * 1e-3) / 1e9,
                                      "what": "the same H2D and D2H byte counts per picture as plain pinned copies on two streams, no kernels: the host / PCIe limit the end-to-end number can approach"}
        # and with the reference's dense int32 coefficient layout going up (what round-1's first number measured)
        ctx.recon_frame_host(f_out, f_refs, descs_dense)
        t0 = time.perf_counter()
        for _ in range(e_steps):
            ctx.recon_frame_host(f_out, f_refs, descs_dense)
        barrier()
        d_ms = (time.perf_counter() - t0) * 1e3
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([d_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            d_ms = float(t.item())
        e2e["dense_int32_layout"] = {"value": luma_px_per_step * e_steps * world / (d_ms * 1e-3) / 1e6, "unit": UNIT,
                                     "h2d_bytes_per_step": int(h2d + h2d_dense_extra), "d2h_bytes_per_step": int(d2h)}
        for t_ in h_out:
            t_.zero_()
        ctx.recon_frame_host_async(f_out, f_refs, descs)    # the timed form: two calls queued back to back, one wait
        ctx.recon_frame_host_async(f_out, f_refs, descs)
        ctx.sync()
        # sanity: the host path produced the same pictures as the device-resident path
        got = out.to_numpy()
        for c in range(3):
            wv = g1.plane_wh(c)[0]
            he, de = h_out[c].numpy().view(np.uint16)[:, :, :wv], got[c][:, :, :wv]
            if not np.array_equal(he, de):
                bad = np.argwhere(he != de)
                per_pic = [int((he[k_] != de[k_]).sum()) for k_ in range(he.shape[0])]
                raise AssertionError("e2e result differs from device-resident result: plane %d, %d samples (per picture %s), first (pic %d, y %d, x %d) %d vs %d" % (
                    c, len(bad), per_pic, bad[0][0], bad[0][1], bad[0][2], he[tuple(bad[0])], de[tuple(bad[0])]))

    # ---- parity gate at the benchmark's own size: the reference C (oracle/_ref, else the oracle port) reconstructs the
    # ---- two distinct pictures of this rank's workload; every ring slot of the device-resident result, and of the
    # ---- end-to-end result, must equal the picture of its content bit for bit ---------------------------
    parity = None
    if not args.no_parity:
        kind_p, fns = cpu_lib()
        want = [None] * inp.distinct

        def ref_pic(i):
            want[i] = [q.copy() for q in cpu_reconstruct(fns, inp, i, inp.ref_planes, [abi.alloc_planes(g1) for _ in range(3)])]

        ts = [threading.Thread(target=ref_pic, args=(i,)) for i in range(inp.distinct)]
        for t in ts: t.start()
        for t in ts: t.join()
        got = out.to_numpy()
        bad = []
        for c in range(3):
            wv = g1.plane_wh(c)[0]
            for k_ in range(frames):
                ref_k = want[k_ % inp.distinct][c][0][:, :wv]
                if not np.array_equal(got[c][k_][:, :wv], ref_k):
                    bad.append(("device", c, k_, int((got[c][k_][:, :wv] != ref_k).sum())))
                if e2e is not None and not np.array_equal(h_out[c].numpy().view(np.uint16)[k_][:, :wv], ref_k):
                    bad.append(("e2e", c, k_, int((h_out[c].numpy().view(np.uint16)[k_][:, :wv] != ref_k).sum())))
        parity = {"pictures": frames, "distinct": inp.distinct, "equal": not bad, "against": kind_p,
                  "paths": ["device-resident"] + (["e2e host entry"] if e2e is not None else [])}
        if bad:
            sys.stderr.write("bench.py: PARITY FAILURE vs the %s C at %dx%d: (path, plane, ring slot, samples) %s\n" % (
                kind_p, args.width, args.height, bad[:8]))
            print(json.dumps({"metric": METRIC, "parity": parity, "error": "result differs from the reference C"}))
            return 3

    # ---- CPU baseline beside it (rank 0, N == 1 only) ------------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        kind, mpix, ms = run_cpu(inp, steps=1, warmup=1, threads=threads)
        _, mpix1, ms1 = run_cpu(inp, 1, 0, 1)
        cpu_baseline = {"value": mpix, "unit": UNIT, "cores": threads, "kind": kind,
                        "sample": "1 warm-up + 1 timed step x %d pictures (one %dx%d picture per host thread, all stages), same synthetic workload, %.0f ms" % (
                            threads, args.width, args.height, ms),
                        "one_thread": {"value": mpix1, "unit": UNIT, "ms_per_picture": ms1}, "cpu_model": cpu_model(),
                        "simd": "C only: the reference's x86 assembly needs nasm, which this image lacks, so its AVX2 MC/SAO/ALF is not represented"}

    if rank == 0:
        ring_mb = (3 * inp.frame_bytes() * frames + sum(len(c) for c in inp.coeffs) * 4 / inp.distinct * frames) / 1e6
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm_steps,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u16", "data": "synthetic",
            "config": config,
            "run": {"pictures_per_step": frames, "pictures_per_launch": group, "ring_mb": ring_mb,
                    "dpb": ("device-resident DPB ring pre-padded with %d replicated samples per side (vvc_cuda_pad_frame, VVC_CUDA_OPT_REF_PAD; "
                            "host reference pictures of the e2e figure are staged without margins)" % args.ref_pad) if args.ref_pad else "device-resident DPB ring, no margins",
                    "l2": "per step %d reference + %d reconstructed + %d output pictures and their coefficients = %.0f MB (> 126 MB)" % (frames, frames, frames, ring_mb)},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(gpu_launches),
            "clocks": clocks, "parity": parity,
        }
        line.update(extra)
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
