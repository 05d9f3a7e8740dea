"""GPU parity: the whole-picture reconstruction entries (INTER -> residual -> inverse LMCS -> deblock V/H -> SAO ->
ALF) - vvc_cuda_recon_frame on a device-resident ring and the pipelined vvc_cuda_recon_frame_host with PINNED host
buffers (so its copies really overlap its kernels) - against the oracle's stage-by-stage chain."""
import ctypes as C

import numpy as np
import pytest

from ffvvc_b200 import abi, synth
from tests import util
from tests.golden_cases import STRESS_MIX

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import torch
    from ffvvc_b200 import lib
    c = lib.Context(0)
    with torch.cuda.stream(c.torch_stream()):
        yield c
    c.close()


def build(w, h, batch, seed, coeff_mode="dense", lmcs_chroma=False):
    g1 = abi.FrameGeom(w, h)
    gr = abi.FrameGeom(w, h, batch=batch)
    case = dict(g1=g1, gr=gr, refs=synth.struct_planes(gr, seed=seed))
    case["pbs"], case["wp"], case["prof"] = synth.pb_list(gr, n_refs=batch, seed=seed + 1, mix=STRESS_MIX)
    case["tbs"], case["coeffs"] = synth.tb_list(gr, seed=seed + 2, lfnst_set_of=util.oracle().vvco_lfnst_tr_set, extras=False)
    case["fmt"], case["quant"], case["sl"] = abi.COEFF_DENSE32, None, None
    if coeff_mode == "window_q":        # quantised levels in the 16-bit window layout, dequantised on the device
        case["tbs"], case["coeffs"] = abi.pack_window16(synth.tb_for_window(case["tbs"]), case["coeffs"])
        case["quant"], case["sl"] = synth.tb_quant(case["tbs"], seed=seed + 7, scaling=True)
        case["fmt"] = abi.COEFF_WINDOW16
    case["vpdus"] = None
    if lmcs_chroma:                     # ph_chroma_residual_scale_flag: luma blocks first, chroma blocks name their VPDU
        case["tbs"], order, case["vpdus"], case["lmcs_params"], case["n_luma"] = synth.lmcs_chroma(gr, case["tbs"], seed=seed + 8)
        if case["quant"] is not None:
            case["quant"] = case["quant"][order].copy()
    _, case["inv"] = synth.lmcs_luts(10, seed=seed + 3)
    case["maps"] = synth.deblock_maps(gr, seed=seed + 4, qp_base=27, qp_span=16)
    case["sao"] = synth.sao_params(gr, seed=seed + 5)
    case["alf"], case["sets"] = synth.alf_params(gr, seed=seed + 6)
    return case


def oracle_chain(case):
    o, gr = util.oracle(), case["gr"]
    cur = abi.alloc_planes(gr, fill=0)
    o.vvco_inter_frame(abi.frame_from_numpy(gr, cur), abi.frame_from_numpy(gr, case["refs"]), case["pbs"].ctypes.data, len(case["pbs"]),
                       case["wp"].ctypes.data, case["prof"].ctypes.data, None)
    co = case["coeffs"].copy()
    cd = abi.coeffs_desc(co.ctypes.data, co.size, case["fmt"], case["quant"].ctypes.data if case["quant"] is not None else None,
                         case["sl"].ctypes.data if case["sl"] is not None else None)
    if case["vpdus"] is not None:
        nl, vp = case["n_luma"], case["vpdus"]
        scales = np.zeros(len(vp), np.uint16)
        o.vvco_itx_frame_q(abi.frame_from_numpy(gr, cur), C.byref(cd), case["tbs"].ctypes.data, nl, 15)
        o.vvco_lmcs_chroma_scale(abi.frame_from_numpy(gr, cur), vp.ctypes.data, len(vp), case["lmcs_params"].ctypes.data, scales.ctypes.data)
        cd.lmcs_scales = scales.ctypes.data
        if case["quant"] is not None:
            cd.quant = case["quant"][nl:].ctypes.data
        o.vvco_itx_frame_q(abi.frame_from_numpy(gr, cur), C.byref(cd), case["tbs"][nl:].ctypes.data, len(case["tbs"]) - nl, 15)
    else:
        o.vvco_itx_frame_q(abi.frame_from_numpy(gr, cur), C.byref(cd), case["tbs"].ctypes.data, len(case["tbs"]), 15)
    o.vvco_lmcs_frame(abi.frame_from_numpy(gr, cur), case["inv"].ctypes.data, None)
    md = abi.deblock_maps_desc(gr, case["maps"])
    a, b = abi.alloc_planes(gr), abi.alloc_planes(gr)
    o.vvco_deblock_frame(abi.frame_from_numpy(gr, a), abi.frame_from_numpy(gr, cur), C.byref(md), 1)
    o.vvco_deblock_frame(abi.frame_from_numpy(gr, b), abi.frame_from_numpy(gr, a), C.byref(md), 0)
    o.vvco_sao_frame(abi.frame_from_numpy(gr, a), abi.frame_from_numpy(gr, b), case["sao"].ctypes.data)
    o.vvco_alf_frame(abi.frame_from_numpy(gr, b), abi.frame_from_numpy(gr, a), case["alf"].ctypes.data, case["sets"].ctypes.data, 0)
    return b


@pytest.mark.parametrize("w,h,batch,seed,coeff_mode,lmcs_chroma", [(416, 240, 4, 31, "dense", False), (256, 192, 5, 32, "dense", False),
                                                                   (416, 240, 3, 33, "window_q", False), (416, 240, 3, 34, "window_q", True),
                                                                   (256, 192, 2, 35, "dense", True)])
def test_recon_entries_bit_exact(ctx, w, h, batch, seed, coeff_mode, lmcs_chroma):
    import torch
    from ffvvc_b200 import device
    case = build(w, h, batch, seed, coeff_mode, lmcs_chroma)
    g1, gr = case["g1"], case["gr"]
    want = oracle_chain(case)
    keep = []

    def up(a):
        t, p = device.to_device(a)
        keep.append(t)
        return p

    # ---- device-resident ring ----
    refs, cur, out = device.DeviceFrames(gr, planes=case["refs"]), device.DeviceFrames(gr), device.DeviceFrames(gr)
    md = abi.deblock_maps_desc(gr, case["maps"], ptr_of=up)
    d = abi.VVCCudaReconDesc()
    d.pbs, d.n_pbs, d.wp, d.n_wp, d.prof, d.n_prof = up(case["pbs"]), len(case["pbs"]), up(case["wp"]), len(case["wp"]), up(case["prof"]), len(case["prof"])
    d.log2_transform_range = 15
    d.coeffs, d.n_coeffs, d.tbs, d.n_tbs = up(case["coeffs"]), len(case["coeffs"]), up(case["tbs"]), len(case["tbs"])
    d.coeff_format = case["fmt"]
    if case["quant"] is not None:
        d.quant, d.scaling = up(case["quant"]), up(case["sl"])
    d.lmcs_inv_lut = up(case["inv"])
    if lmcs_chroma:
        d.lmcs_vpdus, d.n_lmcs_vpdus, d.lmcs_params, d.n_luma_tbs = up(case["vpdus"]), len(case["vpdus"]), up(case["lmcs_params"]), case["n_luma"]
    d.inloop.deblock = C.pointer(md)
    d.inloop.sao, d.inloop.alf, d.inloop.alf_sets = up(case["sao"]), up(case["alf"]), up(case["sets"])
    for _ in range(2):                      # in place: a second pass over the same ring must give the same pictures
        ctx.recon_frame(out.desc, cur.desc, refs.desc, d)
    ctx.sync()
    util.assert_planes_equal(gr, out.to_numpy(), want, "recon_frame vs oracle chain")

    # ---- pipelined host entry, one descriptor set per picture, everything in pinned memory ----
    pins = []

    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).pin_memory()
        pins.append(t)
        return t.data_ptr()

    h_refs = [torch.from_numpy(p.view(np.int16)).pin_memory() for p in case["refs"]]
    h_out = [torch.zeros_like(t).pin_memory() for t in h_refs]
    f_refs = abi.frame_desc(gr, [t.data_ptr() for t in h_refs], [t.stride(1) * 2 for t in h_refs], [t.stride(0) * 2 for t in h_refs])
    f_out = abi.frame_desc(gr, [t.data_ptr() for t in h_out], [t.stride(1) * 2 for t in h_out], [t.stride(0) * 2 for t in h_out])
    descs = (abi.VVCCudaReconDesc * batch)()
    hold = []
    n_ctb = g1.ctb_count
    for k in range(batch):
        pb = case["pbs"][case["pbs"]["pic"] == k].copy()
        pb["pic"] = 0
        of_k = case["tbs"]["pic"] == k
        tb = case["tbs"][of_k].copy()
        tb["pic"] = 0
        lo = int(tb["coeff_offset"].min())
        if case["fmt"] == abi.COEFF_WINDOW16:
            area = tb["nzw"].astype(np.int64) * tb["nzh"].astype(np.int64)
        else:
            area = (1 << tb["log2_w"].astype(np.int64)) * (1 << tb["log2_h"].astype(np.int64))
        hi = int((tb["coeff_offset"].astype(np.int64) + area).max())
        tb["coeff_offset"] -= lo
        hmd = abi.VVCCudaDeblockMaps()
        for dr in range(2):
            for c in range(3):
                rows, pitch = abi.deblock_map_shape(g1, dr, c)
                hmd.edge[dr][c] = pin(case["maps"][dr][c][k])
                hmd.pitch[dr][c], hmd.rows[dr][c], hmd.size[dr][c] = pitch, rows, rows * pitch
        hold.append(hmd)
        e = descs[k]
        e.pbs, e.n_pbs, e.wp, e.n_wp, e.prof, e.n_prof = pin(pb), len(pb), pin(case["wp"]), len(case["wp"]), pin(case["prof"]), len(case["prof"])
        e.log2_transform_range = 15
        e.coeffs, e.n_coeffs, e.tbs, e.n_tbs = pin(case["coeffs"][lo:hi]), hi - lo, pin(tb), len(tb)
        e.coeff_format = case["fmt"]
        if coeff_mode == "window_q":      # name the DPB slots the picture reads: references go up lazily, picture by picture
            gpm = (pb["flags"] & abi.PB_GPM) != 0
            used = np.concatenate([pb["ref"][((pb["pred_flag"] & 1) != 0) | gpm, 0], pb["ref"][((pb["pred_flag"] & 2) != 0) | gpm, 1]])
            e.ref_slots = int(np.bitwise_or.reduce(1 << np.unique(used).astype(np.int64))) if k else 0
        if case["quant"] is not None:
            e.quant, e.scaling = pin(case["quant"][of_k]), pin(case["sl"])
        e.lmcs_inv_lut = pin(case["inv"])
        if lmcs_chroma:                  # this picture's VPDUs; its chroma blocks index them from 1
            vk = case["vpdus"]["pic"] == k
            first = int(np.nonzero(vk)[0][0])
            vp = case["vpdus"][vk].copy()
            vp["pic"] = 0
            tb["chroma_scale"] = np.where(tb["chroma_scale"] > 0, tb["chroma_scale"] - first, 0)
            C.memmove(e.tbs, tb.ctypes.data, tb.nbytes)
            e.lmcs_vpdus, e.n_lmcs_vpdus, e.lmcs_params, e.n_luma_tbs = pin(vp), len(vp), pin(case["lmcs_params"]), int((tb["c_idx"] == 0).sum())
        e.inloop.deblock = C.pointer(hmd)
        e.inloop.sao = pin(case["sao"][k * n_ctb:(k + 1) * n_ctb])
        e.inloop.alf = pin(case["alf"][k * n_ctb:(k + 1) * n_ctb])
        e.inloop.alf_sets = pin(case["sets"])
    ctx.recon_frame_host(f_out, f_refs, descs)
    got = [t.numpy().view(np.uint16) for t in h_out]
    util.assert_planes_equal(gr, got, want, "recon_frame_host (pinned, pipelined) vs oracle chain")
    # the same entry with the DPB resident in HBM (device pointers for refs)
    for t in h_out:
        t.zero_()
    ctx.recon_frame_host(f_out, refs.desc, descs)
    got = [t.numpy().view(np.uint16) for t in h_out]
    util.assert_planes_equal(gr, got, want, "recon_frame_host with a device-resident DPB vs oracle chain")
    # the asynchronous form, three calls back to back (slots and events rotate across the calls, the second and third
    # calls' uploads run under the copy-out of the one before); the completion report arrives after the last copy-out
    for t in h_out:
        t.zero_()
    import threading
    done = threading.Event()
    status = []
    for _ in range(3):
        ctx.recon_frame_host_async(f_out, f_refs, descs)
    keep = ctx.notify(lambda st: (status.append(st), done.set()))
    assert done.wait(60), "no completion report"
    assert status == [0]
    got = [t.numpy().view(np.uint16).copy() for t in h_out]
    ctx.sync()
    util.assert_planes_equal(gr, got, want, "recon_frame_host_async x3 + notify vs oracle chain")
    del keep


def test_recon_4k_bit_exact(ctx):
    """The headline size: two 3840x2160 pictures through vvc_cuda_recon_frame (quantised levels, dequant() on the device,
    stress mix of record kinds) against the oracle's stage-by-stage chain, every sample."""
    from ffvvc_b200 import device
    case = build(3840, 2160, 2, 41, "window_q")
    gr = case["gr"]
    want = oracle_chain(case)
    keep = []

    def up(a):
        t, p = device.to_device(a)
        keep.append(t)
        return p

    refs, cur, out = device.DeviceFrames(gr, planes=case["refs"]), device.DeviceFrames(gr), device.DeviceFrames(gr)
    md = abi.deblock_maps_desc(gr, case["maps"], ptr_of=up)
    d = abi.VVCCudaReconDesc()
    d.pbs, d.n_pbs, d.wp, d.n_wp, d.prof, d.n_prof = up(case["pbs"]), len(case["pbs"]), up(case["wp"]), len(case["wp"]), up(case["prof"]), len(case["prof"])
    d.log2_transform_range = 15
    d.coeffs, d.n_coeffs, d.tbs, d.n_tbs = up(case["coeffs"]), len(case["coeffs"]), up(case["tbs"]), len(case["tbs"])
    d.coeff_format, d.quant, d.scaling = case["fmt"], up(case["quant"]), up(case["sl"])
    d.lmcs_inv_lut = up(case["inv"])
    d.inloop.deblock = C.pointer(md)
    d.inloop.sao, d.inloop.alf, d.inloop.alf_sets = up(case["sao"]), up(case["alf"]), up(case["sets"])
    ctx.recon_frame(out.desc, cur.desc, refs.desc, d)
    ctx.sync()
    util.assert_planes_equal(gr, out.to_numpy(), want, "recon_frame 4K vs oracle chain")


def test_recon_host_entry_arena_and_device_output(ctx):
    """One pinned arena per picture (vvc_cuda_recon_arena_bind: a single upload per picture) and an output ring that stays
    in HBM (device `out`): same pictures as the oracle chain."""
    import torch
    from ffvvc_b200 import device, lib
    batch = 3
    case = build(416, 240, batch, 51, "window_q")
    g1, gr = case["g1"], case["gr"]
    want = oracle_chain(case)
    n_ctb = g1.ctb_count
    keep = []

    def alloc(n):
        t = torch.empty(n + 256, dtype=torch.uint8).pin_memory()
        keep.append(t)
        return t, (t.data_ptr() + 255) & ~255

    descs = (abi.VVCCudaReconDesc * batch)()
    for k in range(batch):
        pb = case["pbs"][case["pbs"]["pic"] == k].copy()
        pb["pic"] = 0
        of_k = case["tbs"]["pic"] == k
        tb = case["tbs"][of_k].copy()
        tb["pic"] = 0
        lo = int(tb["coeff_offset"].min())
        hi = int((tb["coeff_offset"].astype(np.int64) + tb["nzw"].astype(np.int64) * tb["nzh"].astype(np.int64)).max())
        tb["coeff_offset"] -= lo
        d, m, a = abi.recon_arena(lib.load(), g1, alloc, pbs=pb, wp=case["wp"], prof=case["prof"], tbs=tb, coeffs=case["coeffs"][lo:hi],
                                  coeff_format=case["fmt"], quant=case["quant"][of_k], scaling=case["sl"], inv_lut=case["inv"],
                                  maps=[[case["maps"][dr][c][k] for c in range(3)] for dr in range(2)],
                                  sao=case["sao"][k * n_ctb:(k + 1) * n_ctb], alf=case["alf"][k * n_ctb:(k + 1) * n_ctb], sets=case["sets"])
        keep += [m, a]
        C.memmove(C.byref(descs[k]), C.byref(d), C.sizeof(d))
        descs[k].inloop.deblock = C.pointer(m)
    h_refs = [torch.from_numpy(p.view(np.int16)).pin_memory() for p in case["refs"]]
    h_out = [torch.zeros_like(t).pin_memory() for t in h_refs]
    f_refs = abi.frame_desc(gr, [t.data_ptr() for t in h_refs], [t.stride(1) * 2 for t in h_refs], [t.stride(0) * 2 for t in h_refs])
    f_out = abi.frame_desc(gr, [t.data_ptr() for t in h_out], [t.stride(1) * 2 for t in h_out], [t.stride(0) * 2 for t in h_out])
    ctx.recon_frame_host(f_out, f_refs, descs)
    util.assert_planes_equal(gr, [t.numpy().view(np.uint16) for t in h_out], want, "arena-bound descriptors vs oracle chain")
    d_out = device.DeviceFrames(gr)
    ctx.recon_frame_host(d_out.desc, f_refs, descs)          # the output ring stays on the device
    ctx.sync()
    util.assert_planes_equal(gr, d_out.to_numpy(), want, "device-resident output ring vs oracle chain")
