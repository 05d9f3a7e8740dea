#!/usr/bin/env python
"""bench.py - headline benchmark: 4K 10-bit VVC reconstruction + in-loop filtering, Mpix/s per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over a ring of distinct synthetic 4K 10-bit 4:2:0 pictures
that is larger than the 126 MB L2.  Ours: every stage is a CUDA kernel of libvvcdsp_cuda.so called
through the C ABI; `value` has pictures and descriptors resident in HBM, `e2e` goes through the
*_host entry with pinned host buffers (H2D + kernels + D2H inside the timed region).
Reference arm (--impl reference): the reference's own C table entries (oracle/_ref, else the
oracle port) on the host cores, bounded sample, same metric.
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

from ffvvc_b200 import abi, synth  # noqa: E402

METRIC = "vvc_4k10_recon_loopfilter_mpix_per_s"
UNIT = "Mpix/s"
L2_BYTES = 126 * 1024 * 1024


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------
# Synthetic workload (SURVEY.md 8(d) config 2 distributions on 4K pictures)
# ---------------------------------------------------------------------------------------------
class Inputs:
    """Host-side synthetic inputs for a ring of `frames` pictures (distinct seeds per picture)."""

    def __init__(self, width, height, frames, seed=12345, distinct=4):
        self.geom = abi.FrameGeom(width, height, batch=frames)
        distinct = min(distinct, frames)
        g1 = abi.FrameGeom(width, height, batch=distinct)
        base = synth.struct_planes(g1, seed=seed)
        reps = (frames + distinct - 1) // distinct
        # ring slots beyond `distinct` repeat content at different addresses (still defeats L2)
        self.planes = [np.ascontiguousarray(np.concatenate([p] * reps)[:frames]) for p in base]
        m1 = synth.deblock_maps(g1, seed=seed + 1, qp_base=27, qp_span=16)
        self.maps = [[np.ascontiguousarray(np.concatenate([m1[d][c]] * reps)[:frames]) for c in range(3)] for d in range(2)]
        sao1 = synth.sao_params(g1, seed=seed + 2)
        alf1, self.sets = synth.alf_params(g1, seed=seed + 3)
        n1 = g1.ctb_count
        self.sao = np.ascontiguousarray(np.concatenate([sao1] * reps)[:frames * n1])
        self.alf = np.ascontiguousarray(np.concatenate([alf1] * reps)[:frames * n1])

    def desc_bytes(self):
        return sum(m.nbytes for d in self.maps for m in d) + self.sao.nbytes + self.alf.nbytes + self.sets.nbytes

    def frame_bytes(self):
        g = self.geom
        return sum(g.plane_wh(c)[0] * g.plane_wh(c)[1] * 2 for c in range(3)) * g.batch


# algorithmic bytes per luma pixel: one read + one write of all planes per sweep
ALGO_BYTES_PER_LUMA_PX = {"deblock_v": 6.0, "deblock_h": 6.0, "sao": 6.0, "alf": 6.0}
CHAIN_ALGO_BYTES_PER_LUMA_PX = 18.0    # deblock (V+H counted once) + SAO + ALF


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
            time.sleep(0.02)

    def result(self):
        self.stop_flag = True
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ---------------------------------------------------------------------------------------------
# CPU reference arm
# ---------------------------------------------------------------------------------------------
def cpu_chain_lib():
    """Returns (kind, chain(frame_in_planes, maps, sao, alf, sets, geom) -> planes)."""
    FP = C.POINTER(abi.VVCCudaFrame)
    MP = C.POINTER(abi.VVCCudaDeblockMaps)
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libvvcref.so")
    if os.path.exists(ref_so):
        lib, kind, pre = C.CDLL(ref_so), "reference", "vvcref_"
    else:
        lib, kind, pre = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so")), "port", "vvco_"
    dbk = getattr(lib, pre + "deblock_frame")
    sao = getattr(lib, pre + "sao_frame")
    alf = getattr(lib, pre + "alf_frame")
    dbk.argtypes, dbk.restype = [FP, FP, MP, C.c_int], None
    sao.argtypes, sao.restype = [FP, FP, C.c_void_p], None
    alf.argtypes, alf.restype = [FP, FP, C.c_void_p, C.c_void_p, C.c_int], None

    def chain(geom, planes, maps, sao_p, alf_p, sets):
        md = abi.deblock_maps_desc(geom, maps)
        a, b = abi.alloc_planes(geom), abi.alloc_planes(geom)
        fa, fb = abi.frame_from_numpy(geom, a), abi.frame_from_numpy(geom, b)
        dbk(fa, abi.frame_from_numpy(geom, planes), C.byref(md), 1)
        dbk(fb, fa, C.byref(md), 0)
        sao(fa, fb, sao_p.ctypes.data)
        alf(fb, fa, alf_p.ctypes.data, sets.ctypes.data, 0)
        return b

    return kind, chain


def run_cpu_reference(width, height, steps, warmup, threads):
    """Each step: `threads` host threads each push one picture through the reference C chain."""
    kind, chain = cpu_chain_lib()
    inp = Inputs(width, height, frames=1)
    g1 = abi.FrameGeom(width, height)
    work = [(g1, [p.copy() for p in inp.planes], [[m.copy() for m in d] for d in inp.maps], inp.sao.copy(), inp.alf.copy(), inp.sets.copy())
            for _ in range(threads)]

    def one_step():
        ts = [threading.Thread(target=chain, args=w) for w in work]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        return time.perf_counter() - t0

    for _ in range(warmup):
        one_step()
    times = [one_step() for _ in range(steps)]
    total = sum(times)
    mpix = width * height * threads * steps / total / 1e6
    return kind, mpix, total / steps * 1e3


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
    ap.add_argument("--group", type=int, default=1, help="pictures per launch (stage kernels run group by group)")
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    threads = args.cpu_threads or (os.cpu_count() or 1)
    workload = "inloop_4k: deblock V+H -> SAO -> ALF/CC-ALF on %dx%d 10-bit 4:2:0 (MC, itx, LMCS stages not yet in the timed chain)" % (args.width, args.height)

    if args.impl == "reference":
        if rank != 0:
            return 0
        kind, mpix, ms = run_cpu_reference(args.width, args.height, args.steps, max(args.warmup, 1), threads)
        sample = "%d pictures per step (one per host thread) of the same synthetic workload" % threads
        line = {
            "impl": "reference", "metric": METRIC, "value": mpix, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u16", "data": "synthetic", "config": {"workload": workload, "cpu_threads": threads},
            "cpu_baseline": {"value": mpix, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
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
        dist.init_process_group("nccl", device_id=torch.device(dev))

    frames, group = args.frames, max(1, min(args.group, args.frames))
    while frames % group:
        group -= 1
    inp = Inputs(args.width, args.height, frames, seed=12345 + rank)
    geom = inp.geom
    ggeom = abi.FrameGeom(args.width, args.height, batch=group)
    ctx = lib.Context(local_rank)
    stream = ctx.torch_stream()
    torch.cuda.set_stream(stream)      # uploads, events and kernels all on the context's stream

    # ---- device-resident ring -------------------------------------------------------------
    src = device.DeviceFrames(geom, device=dev, planes=inp.planes)
    dst = device.DeviceFrames(geom, device=dev)
    tmp_a = device.DeviceFrames(ggeom, device=dev)
    tmp_b = device.DeviceFrames(ggeom, device=dev)
    keep = []
    map_ptr = [[None] * 3, [None] * 3]
    for d in range(2):
        for c in range(3):
            t, ptr = device.to_device(inp.maps[d][c], dev)
            keep.append(t)
            map_ptr[d][c] = ptr
    t_sao, p_sao = device.to_device(inp.sao, dev)
    t_alf, p_alf = device.to_device(inp.alf, dev)
    t_set, p_set = device.to_device(inp.sets, dev)

    def sub_frame(df, k0):
        f = abi.VVCCudaFrame()
        C.memmove(C.byref(f), C.byref(df.desc), C.sizeof(f))
        for c in range(3):
            f.data[c] = df.desc.data[c] + k0 * df.desc.batch_stride[c]
        f.batch = group
        return f

    groups = []
    n_ctb = geom.ctb_count
    for k0 in range(0, frames, group):
        md = abi.VVCCudaDeblockMaps()
        for d in range(2):
            for c in range(3):
                rows, pitch = abi.deblock_map_shape(geom, d, c)
                md.edge[d][c] = map_ptr[d][c] + k0 * rows * pitch * 4
                md.pitch[d][c], md.rows[d][c], md.size[d][c] = pitch, rows, rows * pitch
        groups.append((sub_frame(src, k0), sub_frame(dst, k0), md,
                       p_sao + k0 * n_ctb * abi.SAO_CTB_DTYPE.itemsize, p_alf + k0 * n_ctb * abi.ALF_CTB_DTYPE.itemsize))

    stage_names = ["deblock_v", "deblock_h", "sao", "alf"]
    launches_per_step = len(groups) * 4

    def step(events=None):
        for gi, (fs, fd, md, ps, pa) in enumerate(groups):
            ev = events[gi] if events is not None else None
            if ev: ev[0].record()
            ctx.deblock_frame(tmp_a.desc, fs, md, 1)
            if ev: ev[1].record()
            ctx.deblock_frame(tmp_b.desc, tmp_a.desc, md, 0)
            if ev: ev[2].record()
            ctx.sao_frame(tmp_a.desc, tmp_b.desc, ps)
            if ev: ev[3].record()
            ctx.alf_frame(fd, tmp_a.desc, pa, p_set, 0)
            if ev: ev[4].record()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- timed region: K steps, device events; per-kernel events ride along --------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ctx.launches
    evs = [[[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in groups] for _ in range(args.steps)]
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

    stage_ms = np.zeros(4)
    for s in range(args.steps):
        for g in evs[s]:
            for i in range(4):
                stage_ms[i] += g[i].elapsed_time(g[i + 1])
    stage_ms /= args.steps * len(groups)          # average duration of one launch of each stage

    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())

    luma_px_per_step = args.width * args.height * frames
    value = luma_px_per_step * args.steps * world / (elapsed_ms * 1e-3) / 1e6

    # ---- roofline of the dominant kernel --------------------------------------------------------
    peak, peak_src = load_peaks()
    dom = int(np.argmax(stage_ms))
    px_per_launch = args.width * args.height * group
    algo_bytes = ALGO_BYTES_PER_LUMA_PX[stage_names[dom]] * px_per_launch
    achieved = algo_bytes / (stage_ms[dom] * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "kernel": stage_names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": algo_bytes,
        "stage_ms_per_launch": {n: float(v) for n, v in zip(stage_names, stage_ms)},
        "chain": {"algorithmic_bytes_per_luma_px": CHAIN_ALGO_BYTES_PER_LUMA_PX,
                  "achieved_gbs": CHAIN_ALGO_BYTES_PER_LUMA_PX * luma_px_per_step * args.steps / (elapsed_ms * 1e-3) / 1e9 / 1.0,
                  "frac": CHAIN_ALGO_BYTES_PER_LUMA_PX * luma_px_per_step * args.steps / (elapsed_ms * 1e-3) / 1e9 / peak},
    }

    # ---- e2e: pinned host pictures through the *_host entry -------------------------------------
    e2e = None
    if not args.no_e2e:
        hp_in = [torch.from_numpy(p.view(np.int16)).pin_memory() for p in inp.planes]
        hp_out = [torch.empty_like(t).pin_memory() for t in hp_in]
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).pin_memory()
        h_maps = [[pin(inp.maps[d][c]) for c in range(3)] for d in range(2)]
        h_sao, h_alf, h_set = pin(inp.sao), pin(inp.alf), pin(inp.sets)
        f_in = abi.frame_desc(geom, [t.data_ptr() for t in hp_in], [t.stride(1) * 2 for t in hp_in], [t.stride(0) * 2 for t in hp_in])
        f_out = abi.frame_desc(geom, [t.data_ptr() for t in hp_out], [t.stride(1) * 2 for t in hp_out], [t.stride(0) * 2 for t in hp_out])
        hmd = abi.VVCCudaDeblockMaps()
        for d in range(2):
            for c in range(3):
                rows, pitch = abi.deblock_map_shape(geom, d, c)
                hmd.edge[d][c] = h_maps[d][c].data_ptr()
                hmd.pitch[d][c], hmd.rows[d][c], hmd.size[d][c] = pitch, rows, rows * pitch
        hdesc = abi.inloop_desc(hmd, h_sao.data_ptr(), h_alf.data_ptr(), h_set.data_ptr())
        e_steps = max(2, min(args.steps, 5))
        for _ in range(2):
            ctx.inloop_frame_host(f_out, f_in, hdesc)
        barrier()
        t0 = time.perf_counter()
        es, ee = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        es.record()
        for _ in range(e_steps):
            ctx.inloop_frame_host(f_out, f_in, hdesc)      # returns after the D2H copy finished
        ee.record()
        barrier()
        e_ms = max(es.elapsed_time(ee), (time.perf_counter() - t0) * 1e3 * 0.0)
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([e_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_ms = float(t.item())
        e2e = {"value": luma_px_per_step * e_steps * world / (e_ms * 1e-3) / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": int(inp.frame_bytes() + inp.desc_bytes()), "d2h_bytes_per_step": int(inp.frame_bytes()),
               "steps": e_steps, "api": "vvc_cuda_inloop_frame_host (pinned host pictures + descriptors)"}
        # sanity: the host path produced the same pictures as the device path
        got = dst.to_numpy()
        for c in range(3):
            assert np.array_equal(hp_out[c].numpy().view(np.uint16), got[c]), "e2e result differs from device-resident result"

    # ---- CPU baseline beside it (rank 0, N == 1 only) --------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        kind, mpix, ms = run_cpu_reference(args.width, args.height, steps=2, warmup=1, threads=threads)
        cpu_baseline = {"value": mpix, "unit": UNIT, "cores": threads, "kind": kind,
                        "sample": "2 timed steps x %d pictures (one %dx%d picture per host thread), same synthetic workload" % (threads, args.width, args.height)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u16", "data": "synthetic",
            "config": {"workload": workload, "pictures_per_step": frames, "pictures_per_launch": group,
                       "l2": "inputs larger than L2: ring of %d pictures = %.0f MB per plane set (> 126 MB), no flush needed" % (frames, inp.frame_bytes() / 1e6),
                       "parallelism": "independent streams x%d, no collective" % world},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(gpu_launches),
            "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
