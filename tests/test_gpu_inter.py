"""GPU parity: inter prediction stage (8/4-tap MC, bi-pred avg / w_avg / BCW, DMVR, BDOF, PROF, GPM,
picture-border and DMVR-window clamps) vs the oracle, through the C ABI."""
import numpy as np
import pytest

from ffvvc_b200 import abi, synth
from tests import util
from tests.test_oracle_vs_ref_inter import STRESS_MIX, make_case, run_inter

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import torch
    from ffvvc_b200 import lib
    c = lib.Context(0)
    with torch.cuda.stream(c.torch_stream()):
        yield c
    c.close()


def cuda_inter(ctx, gd, gr, refs, pbs, wp, prof):
    from ffvvc_b200 import device
    dst = device.DeviceFrames(gd, planes=abi.alloc_planes(gd, fill=77))
    ref = device.DeviceFrames(gr, planes=refs)
    t1, p1 = device.to_device(pbs)
    t2, p2 = device.to_device(wp)
    t3, p3 = device.to_device(prof)
    t4, p4 = device.to_device(np.zeros(len(pbs), dtype=abi.DMVR_OUT_DTYPE))
    ctx.inter_frame(dst.desc, ref.desc, p1, len(pbs), p2, p3, p4)
    ctx.sync()
    return dst.to_numpy(), t4.cpu().numpy().view(abi.DMVR_OUT_DTYPE)


@pytest.mark.parametrize("w,h,seed,uniform,bd", [(416, 240, 1, False, 10), (416, 240, 2, True, 10), (256, 192, 3, False, 10),
                                                  (832, 480, 4, False, 10), (136, 72, 5, True, 10), (256, 128, 9, False, 12),
                                                  (1920, 1080, 6, False, 10)])
def test_inter_stage_bit_exact(ctx, w, h, seed, uniform, bd):
    gd, gr, refs, pbs, wp, prof = make_case(w, h, seed, mix=STRESS_MIX if seed != 6 else None, uniform=uniform, bit_depth=bd)
    gp, go = cuda_inter(ctx, gd, gr, refs, pbs, wp, prof)
    od, oo = run_inter(util.oracle().vvco_inter_frame, gd, gr, refs, pbs, wp, prof)
    util.assert_planes_equal(gd, gp, od, "cuda vs oracle")
    dm = (pbs["flags"] & abi.PB_DMVR) != 0
    assert np.array_equal(go[dm], oo[dm]), "DMVR outputs (refined vectors, min SAD, BDOF decision) differ"


@pytest.mark.parametrize("tma", [0, 1])
@pytest.mark.parametrize("w,h,seed,uniform", [(416, 240, 1, False), (136, 72, 5, True), (832, 480, 4, False), (1920, 1080, 6, False)])
def test_inter_tma_staged_windows_bit_exact(ctx, w, h, seed, uniform, tma):
    """VVC_CUDA_OPT_INTER_TMA on / off: DMVR windows fetched by the copy engine one record ahead (interior records) or
    staged by the warp's own loads (always for records whose windows leave the picture); both against the oracle,
    refined vectors included."""
    gd, gr, refs, pbs, wp, prof = make_case(w, h, seed, mix=STRESS_MIX if seed != 6 else None, uniform=uniform)
    ctx.set_option(abi.OPT_INTER_TMA, tma)
    try:
        gp, go = cuda_inter(ctx, gd, gr, refs, pbs, wp, prof)
    finally:
        ctx.set_option(abi.OPT_INTER_TMA, abi.INTER_TMA_DEFAULT)
    od, oo = run_inter(util.oracle().vvco_inter_frame, gd, gr, refs, pbs, wp, prof)
    util.assert_planes_equal(gd, gp, od, "cuda (tma=%d) vs oracle" % tma)
    dm = (pbs["flags"] & abi.PB_DMVR) != 0
    assert np.array_equal(go[dm], oo[dm])


@pytest.mark.parametrize("pad", [16, 64, 128])
@pytest.mark.parametrize("w,h,seed,uniform", [(416, 240, 1, False), (136, 72, 5, True), (832, 480, 4, False), (1920, 1080, 6, False)])
def test_inter_pre_padded_reference_planes_bit_exact(ctx, w, h, seed, uniform, pad):
    """VVC_CUDA_OPT_REF_PAD: a DPB ring with a margin of replicated samples (vvc_cuda_pad_frame) gives the pictures of the
    per-block edge emulation; vectors that leave the margin still take the clamped path.  The margin itself must be the
    replicated border."""
    from ffvvc_b200 import device
    gd, gr, refs, pbs, wp, prof = make_case(w, h, seed, mix=STRESS_MIX if seed != 6 else None, uniform=uniform)
    dst = device.DeviceFrames(gd, planes=abi.alloc_planes(gd, fill=77))
    ref = device.DeviceFrames(gr, planes=refs, pad=pad)
    ctx.pad_frame(ref.desc, pad)
    full = ref.to_numpy(with_margin=True)
    for c in range(3):
        mx, my = ref.margin[c]
        pw, ph = gr.plane_wh(c)
        ys = np.clip(np.arange(-my, ph + my), 0, ph - 1)
        xs = np.clip(np.arange(-mx, pw + mx), 0, pw - 1)
        assert np.array_equal(full[c][:, :, :pw + 2 * mx], refs[c][:, ys][:, :, xs]), "margin of plane %d" % c
    keep = [device.to_device(a) for a in (pbs, wp, prof, np.zeros(len(pbs), dtype=abi.DMVR_OUT_DTYPE))]
    ctx.set_option(abi.OPT_REF_PAD, pad)
    try:
        ctx.inter_frame(dst.desc, ref.desc, keep[0][1], len(pbs), keep[1][1], keep[2][1], keep[3][1])
        ctx.sync()
    finally:
        ctx.set_option(abi.OPT_REF_PAD, 0)
    od, oo = run_inter(util.oracle().vvco_inter_frame, gd, gr, refs, pbs, wp, prof)
    util.assert_planes_equal(gd, dst.to_numpy(), od, "pre-padded references (margin %d) vs oracle" % pad)
    dm = (pbs["flags"] & abi.PB_DMVR) != 0
    assert np.array_equal(keep[3][0].cpu().numpy().view(abi.DMVR_OUT_DTYPE)[dm], oo[dm])


@pytest.mark.parametrize("pad", [0, 16])
def test_inter_patch_window_alignment_sweep(ctx, pad):
    """The staged windows of the patch kernels: every record shape (4 / 8 / 16 wide and high), uni and bi, with the integer
    part of the horizontal vector walking through every alignment of a window against the 8 / 16-byte requests and across the
    left / right picture border (records that stay inside take the cp.async regions, the others the clamped path), every
    fractional phase; without and with a pre-padded DPB.  Against the oracle, every sample."""
    from ffvvc_b200 import device
    W, H = 256, 128
    gd, gr, refs, _, wp, prof = make_case(W, H, 77)
    shapes = [(w, h) for w in (4, 8, 16) for h in (4, 8, 16) if (w, h) != (4, 4)]
    recs = []
    n = 0
    for cy in range(H // 16):
        for cx in range(W // 16):
            w, h = shapes[(cx + 3 * cy) % len(shapes)]
            bi = w >= 8 and h >= 8 and (cx + cy) % 3 != 0
            for y in range(0, 16, h):
                for x in range(0, 16, w):
                    r = np.zeros(1, dtype=abi.PB_DTYPE)
                    r["x0"], r["y0"], r["w"], r["h"] = cx * 16 + x, cy * 16 + y, w, h
                    r["planes"] = abi.PB_LUMA | abi.PB_CHROMA
                    r["pred_flag"] = abi.PF_BI if bi else (abi.PF_L0 if n % 2 else abi.PF_L1)
                    r["ref"] = [(n // 3) % gr.batch, (n // 5 + 1) % gr.batch]
                    for l in range(2):
                        ix = (n * (3 + 4 * l) + 5 * l) % 23 - 11            # -11 .. 11 samples: every residue of the window start
                        iy = (n * (5 - 2 * l) + l) % 9 - 4
                        r["mv"][0, l, 0] = ix * 16 + (n * (7 + l)) % 16
                        r["mv"][0, l, 1] = iy * 16 + (n * (11 + 2 * l) + 3) % 16
                    if n % 7 == 0:
                        r["mv"][0, :, 0] &= ~15                             # integer columns
                    if n % 11 == 0:
                        r["mv"][0, :, 1] &= ~15                             # integer rows
                    r["filt"] = 1 if n % 13 == 0 else 0
                    recs.append(r)
                    n += 1
    pbs = np.concatenate(recs)
    dst = device.DeviceFrames(gd, planes=abi.alloc_planes(gd, fill=77))
    ref = device.DeviceFrames(gr, planes=refs, pad=pad)
    if pad:
        ctx.pad_frame(ref.desc, pad)
    keep = [device.to_device(a) for a in (pbs, wp, prof, np.zeros(len(pbs), dtype=abi.DMVR_OUT_DTYPE))]
    ctx.set_option(abi.OPT_REF_PAD, pad)
    try:
        ctx.inter_frame(dst.desc, ref.desc, keep[0][1], len(pbs), keep[1][1], keep[2][1], keep[3][1])
        ctx.sync()
    finally:
        ctx.set_option(abi.OPT_REF_PAD, 0)
    od, _ = run_inter(util.oracle().vvco_inter_frame, gd, gr, refs, pbs, wp, prof)
    util.assert_planes_equal(gd, dst.to_numpy(), od, "patch kernels, window alignment sweep (margin %d) vs oracle" % pad)


def test_inter_generic_kernel_matches(ctx):
    """The generic CTA-per-record kernel (any bit depth / alignment) and the 10-bit warp-per-record kernel agree."""
    gd, gr, refs, pbs, wp, prof = make_case(416, 240, 11, mix=STRESS_MIX)
    fast, fo = cuda_inter(ctx, gd, gr, refs, pbs, wp, prof)
    ctx.set_option(abi.OPT_GENERIC_KERNELS, 1)
    try:
        gen, go = cuda_inter(ctx, gd, gr, refs, pbs, wp, prof)
    finally:
        ctx.set_option(abi.OPT_GENERIC_KERNELS, 0)
    util.assert_planes_equal(gd, fast, gen, "warp kernel vs generic kernel")
    assert np.array_equal(fo, go)


def test_inter_ring_and_host_entry(ctx):
    """Two destination pictures predicted from a 3-picture DPB ring; the *_host entry gives the same result."""
    gd, gr, refs, pbs, wp, prof = make_case(256, 128, 21, mix=STRESS_MIX, batch=2)
    gp, go = cuda_inter(ctx, gd, gr, refs, pbs, wp, prof)
    od, oo = run_inter(util.oracle().vvco_inter_frame, gd, gr, refs, pbs, wp, prof)
    util.assert_planes_equal(gd, gp, od, "cuda vs oracle (ring)")
    hd = abi.alloc_planes(gd, fill=77)
    ho = np.zeros(len(pbs), dtype=abi.DMVR_OUT_DTYPE)
    ctx.inter_frame_host(abi.frame_from_numpy(gd, hd), abi.frame_from_numpy(gr, refs), pbs.ctypes.data, len(pbs),
                         wp.ctypes.data, len(wp), prof.ctypes.data, len(prof), ho.ctypes.data)
    util.assert_planes_equal(gd, hd, od, "host entry vs oracle")
    dm = (pbs["flags"] & abi.PB_DMVR) != 0
    assert np.array_equal(ho[dm], oo[dm])


def test_inter_4k_properties(ctx):
    """BASELINE config 4 size: properties that need no CPU pass over the whole picture.
    (a) integer-pel uni prediction with a zero vector copies the reference; (b) the full 4K mix equals the
    oracle on a sample of records re-run in isolation."""
    gd = abi.FrameGeom(3840, 2160)
    gr = abi.FrameGeom(3840, 2160, batch=2)
    refs = synth.struct_planes(abi.FrameGeom(3840, 2160, batch=1), seed=3)
    refs = [np.ascontiguousarray(np.concatenate([p, p[:, ::-1]])) for p in refs]
    pbs, wp, prof = synth.pb_list(gd, n_refs=2, seed=8)
    # (a)
    ys, xs = np.mgrid[0:2160 // 16, 0:3840 // 16]
    cp = np.zeros(ys.size, dtype=abi.PB_DTYPE)
    cp["x0"], cp["y0"], cp["w"], cp["h"] = xs.reshape(-1) * 16, ys.reshape(-1) * 16, 16, 16
    cp["planes"], cp["pred_flag"] = 3, abi.PF_L0
    gp, _ = cuda_inter(ctx, gd, gr, refs, cp, wp, prof)
    for c in range(3):
        wv = gd.plane_wh(c)[0]
        assert np.array_equal(gp[c][0, :, :wv], refs[c][0, :, :wv])
    # (b)
    gp, go = cuda_inter(ctx, gd, gr, refs, pbs, wp, prof)
    pick = synth.LCG(4).below(4000, len(pbs))
    sub = pbs[pick]
    od, oo = run_inter(util.oracle().vvco_inter_frame, gd, gr, refs, sub, wp, prof)
    for r in sub:
        for c in range(3):
            if not (r["planes"] & (2 if c else 1)):
                continue
            sh = 1 if c else 0
            x0, y0, bw, bh = r["x0"] >> sh, r["y0"] >> sh, r["w"] >> sh, r["h"] >> sh
            assert np.array_equal(gp[c][0, y0:y0 + bh, x0:x0 + bw], od[c][0, y0:y0 + bh, x0:x0 + bw]), (r, c)
    dm = (sub["flags"] & abi.PB_DMVR) != 0
    assert np.array_equal(go[pick][dm], oo[dm])


def test_inter_long_launch_matches_short_launches(ctx):
    """A record list above the library's 200 k-record threshold runs its class kernels on one stream, shorter lists on
    parallel streams: the same 4K records as one long launch and as per-picture launches must give the same pictures
    (and one of those pictures is checked against the oracle)."""
    from ffvvc_b200 import device
    batch = 3
    g1 = abi.FrameGeom(3840, 2160)
    gd = abi.FrameGeom(3840, 2160, batch=batch)
    refs = synth.struct_planes(abi.FrameGeom(3840, 2160, batch=2), seed=5)
    pbs1, wp, prof = synth.pb_list(g1, n_refs=2, seed=6)
    parts = []
    for k in range(batch):
        p = pbs1.copy()
        p["pic"] = k
        parts.append(p)
    pbs = np.concatenate(parts)
    assert len(pbs) > 200000 > len(pbs1)
    gr = abi.FrameGeom(3840, 2160, batch=2)
    long_run, _ = cuda_inter(ctx, gd, gr, refs, pbs, wp, prof)
    short_run, _ = cuda_inter(ctx, g1, gr, refs, pbs1, wp, prof)
    for k in range(batch):
        util.assert_planes_equal(g1, [p[k:k + 1] for p in long_run], short_run, "long launch picture %d vs short launch" % k)
    od, _ = run_inter(util.oracle().vvco_inter_frame, g1, gr, refs, pbs1, wp, prof)
    util.assert_planes_equal(g1, short_run, od, "4K picture vs oracle")
