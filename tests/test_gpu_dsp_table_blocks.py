"""GPU parity of the per-call table entries ff_vvc_dsp_init_cuda installs for inter / sao / alf / lf / joint residuals:
every entry is called through the reference's own signature with host buffers, next to the same entry of the compiled
reference's C table (oracle/_ref/libvvcref.so), on the same random inputs; every byte the entries may touch is compared
(the buffers are larger than the blocks, so a write outside the block shows too).  10 and 12 bit."""
import ctypes as C

import numpy as np
import pytest

from ffvvc_b200 import lib, synth
from ffvvc_b200.dsp_tables import VVCDSPContext
from tests import util

pytestmark = pytest.mark.gpu
PB = 128


@pytest.fixture(scope="module", params=[10, 12])
def tables(request):
    if not util.have_ref():
        pytest.skip("compiled reference not built")
    bd = request.param
    ours = VVCDSPContext()
    lib.load().ff_vvc_dsp_cuda_reset_error()
    lib.load().ff_vvc_dsp_init_cuda(C.byref(ours), bd)
    yield bd, ours, util.ref_dsp(bd)
    assert lib.load().ff_vvc_dsp_cuda_last_error() == 0, lib.load().ff_vvc_dsp_cuda_error_string()


class R:
    def __init__(self, seed):
        self.g = synth.LCG(seed)

    def pix(self, n, bd):
        return (self.g.take(n) & ((1 << bd) - 1)).astype(np.uint16)

    def s14(self, n):                      # 14-bit intermediates as the put() entries produce them (signed around 0..16383)
        return ((self.g.take(n) & 0x3fff).astype(np.int64) - (self.g.take(n) & 1) * 8192).astype(np.int16)

    def below(self, n, m):
        return self.g.below(n, m)

    def one(self, m):
        return int(self.g.below(1, m)[0])


def both(fn_o, fn_r, bufs, call):
    """Run `call(fn, *copies of bufs)` for both tables; all buffers must end up identical.  Returns the results."""
    a = [b.copy() for b in bufs]
    b = [x.copy() for x in bufs]
    ra, rb = call(fn_o, *a), call(fn_r, *b)
    for i, (x, y) in enumerate(zip(a, b)):
        assert np.array_equal(x, y), "buffer %d differs: %d values, first at %s" % (i, int((x != y).sum()), np.argwhere(x != y)[0])
    assert ra == rb
    return a


def p8(a, off=0):
    return C.c_void_p(a.ctypes.data + off)


def test_put_family(tables):
    bd, o, r = tables
    rng = R(1)
    lf, cf = synth.spec_table("luma_mc_filters").astype(np.int8), synth.spec_table("chroma_mc_filters").astype(np.int8)
    i8 = C.POINTER(C.c_int8)
    n = 0
    for ch in range(2):
        for (w, h) in [(4, 4), (8, 16), (16, 8), (32, 32), (64, 16), (128, 8)] + ([(2, 2), (2, 8)] if ch else []):
            idx = int(np.log2(w)) - 1
            for vf in range(2):
                for hf in range(2):
                    S = w + 16
                    src = rng.pix((h + 16) * S, bd)
                    fh = np.ascontiguousarray((lf[rng.one(3)][rng.one(15) + 1] if not ch else cf[rng.one(cf.shape[0])][rng.one(31) + 1]))
                    fv = np.ascontiguousarray((lf[rng.one(3)][rng.one(15) + 1] if not ch else cf[rng.one(cf.shape[0])][rng.one(31) + 1]))
                    off = (8 * S + 8) * 2
                    d16 = np.full(h * PB + 8, 77, dtype=np.int16)
                    both(o.inter.put[ch][idx][vf][hf], r.inter.put[ch][idx][vf][hf], [d16, src],
                         lambda f, d, s: f(d.ctypes.data_as(C.POINTER(C.c_int16)), p8(s, off), S * 2, h, fh.ctypes.data_as(i8), fv.ctypes.data_as(i8), w))
                    dp = np.full((h + 1) * (w + 8), 5, dtype=np.uint16)
                    both(o.inter.put_uni[ch][idx][vf][hf], r.inter.put_uni[ch][idx][vf][hf], [dp, src],
                         lambda f, d, s: f(p8(d), (w + 8) * 2, p8(s, off), S * 2, h, fh.ctypes.data_as(i8), fv.ctypes.data_as(i8), w))
                    den, wx, ox = rng.one(8), rng.one(256) - 128, rng.one(256) - 128
                    both(o.inter.put_uni_w[ch][idx][vf][hf], r.inter.put_uni_w[ch][idx][vf][hf], [dp, src],
                         lambda f, d, s: f(p8(d), (w + 8) * 2, p8(s, off), S * 2, h, den, wx, ox, fh.ctypes.data_as(i8), fv.ctypes.data_as(i8), w))
                    n += 3
    assert n > 100


def test_blends(tables):
    bd, o, r = tables
    rng = R(2)
    i16 = C.POINTER(C.c_int16)
    for (w, h) in [(2, 2), (4, 8), (16, 16), (64, 32), (128, 128), (8, 4)]:
        s0, s1 = rng.s14(h * PB), rng.s14(h * PB)
        dp = np.full((h + 1) * (w + 4), 9, dtype=np.uint16)
        both(o.inter.avg, r.inter.avg, [dp, s0, s1], lambda f, d, a, b: f(p8(d), (w + 4) * 2, a.ctypes.data_as(i16), b.ctypes.data_as(i16), w, h))
        den, w0, w1, o0, o1 = rng.one(8), rng.one(256) - 128, rng.one(256) - 128, rng.one(256) - 128, rng.one(256) - 128
        both(o.inter.w_avg, r.inter.w_avg, [dp, s0, s1],
             lambda f, d, a, b: f(p8(d), (w + 4) * 2, a.ctypes.data_as(i16), b.ctypes.data_as(i16), w, h, den, w0, w1, o0, o1))
        if w >= 8 and h >= 8 and w <= 64 and h <= 64:
            wt = rng.below(112 * 112, 9).astype(np.uint8)
            for (sx, sy, org) in [(1, 112, 0), (-1, 112, 111), (112, 1, 0), (1, -112, 111 * 112)]:
                both(o.inter.put_gpm, r.inter.put_gpm, [dp, s0, s1, wt],
                     lambda f, d, a, b, g: f(p8(d), (w + 4) * 2, w, h, a.ctypes.data_as(i16), b.ctypes.data_as(i16), p8(g, org), sx, sy))
        inter = rng.pix(h * (w + 2), bd)
        dp2 = rng.pix((h + 1) * (w + 4), bd)
        for iw in (1, 2, 3):
            both(o.inter.put_ciip, r.inter.put_ciip, [dp2, inter], lambda f, d, s: f(p8(d), (w + 4) * 2, w, h, p8(s), (w + 2) * 2, iw))


def test_dmvr_sad_prof_bdof(tables):
    bd, o, r = tables
    rng = R(3)
    i16 = C.POINTER(C.c_int16)
    for (w, h) in [(8, 16), (16, 8), (16, 16)]:
        S = w + 12
        src = rng.pix((h + 8) * S, bd)
        for vf in range(2):
            for hf in range(2):
                d16 = np.full((h + 5) * PB, 3, dtype=np.int16)
                mx, my = rng.one(15) + 1, rng.one(15) + 1
                both(o.inter.dmvr[vf][hf], r.inter.dmvr[vf][hf], [d16, src], lambda f, d, s: f(p8(d), p8(s, (S + 1) * 2), S * 2, h + 4, mx, my, w + 4))
        t0, t1 = (rng.pix((h + 4) * PB, 10)).astype(np.int16), (rng.pix((h + 4) * PB, 10)).astype(np.int16)
        for dx in range(5):
            for dy in range(5):
                both(o.inter.sad, r.inter.sad, [t0, t1], lambda f, a, b: f(p8(a), p8(b), dx, dy, w, h))
        # BDOF: tiles with a ring (origin at row 1, column 1 of the staging tile)
        a0, a1 = rng.s14((h + 2) * PB + 2), rng.s14((h + 2) * PB + 2)
        dp = np.full((h + 1) * (w + 4), 9, dtype=np.uint16)
        org = (PB + 1) * 2
        both(o.inter.apply_bdof, r.inter.apply_bdof, [dp, a0, a1], lambda f, d, x, y: f(p8(d), (w + 4) * 2, p8(x, org), p8(y, org), w, h))
        srcp = rng.pix((h + 6) * S, bd)
        tile = rng.s14((h + 3) * PB)
        for (xf, yf) in [(0, 0), (5, 9), (8, 15), (15, 3)]:
            both(o.inter.bdof_fetch_samples, r.inter.bdof_fetch_samples, [tile, srcp],
                 lambda f, d, s: f(p8(d, org), p8(s, (2 * S + 2) * 2), S * 2, xf, yf, w, h))
            both(o.inter.fetch_samples, r.inter.fetch_samples, [tile, srcp], lambda f, d, s: f(p8(d, org), p8(s, (2 * S + 2) * 2), S * 2, xf, yf))
        for pad in (0, 1):
            gs = w + 2 + 3
            gh, gv = np.full((h + 3) * gs, 11, dtype=np.int16), np.full((h + 3) * gs, 12, dtype=np.int16)
            both(o.inter.prof_grad_filter, r.inter.prof_grad_filter, [gh, gv, a0], lambda f, a, b, s: f(p8(a), p8(b), gs, p8(s, org), PB, w, h, pad))
    t = rng.s14(8 * PB)
    dmx, dmy = (rng.below(16, 64) - 32).astype(np.int16), (rng.below(16, 64) - 32).astype(np.int16)
    d16 = np.full(5 * PB, 1, dtype=np.int16)
    org = (PB + 1) * 2
    both(o.inter.apply_prof, r.inter.apply_prof, [d16, t], lambda f, d, s: f(p8(d), p8(s, org), dmx.ctypes.data_as(i16), dmy.ctypes.data_as(i16)))
    dp = np.full(5 * 12, 2, dtype=np.uint16)
    both(o.inter.apply_prof_uni, r.inter.apply_prof_uni, [dp, t], lambda f, d, s: f(p8(d), 24, p8(s, org), dmx.ctypes.data_as(i16), dmy.ctypes.data_as(i16)))
    both(o.inter.apply_prof_uni_w, r.inter.apply_prof_uni_w, [dp, t],
         lambda f, d, s: f(p8(d), 24, p8(s, org), dmx.ctypes.data_as(i16), dmy.ctypes.data_as(i16), 3, -57, 21))


def test_sao_entries(tables):
    bd, o, r = tables
    rng = R(4)
    i16 = C.POINTER(C.c_int16)
    S = (2 * PB + 64) // 2                 # the edge entry's implicit source stride: (2 * MAX_PB_SIZE + 64) BYTES, i.e. 160 samples
    for i, size in enumerate([8, 16, 32, 48, 64, 80, 96, 112, 128]):
        for w in (size, size - 4):
            src = rng.pix((size + 2) * S + 64, bd)
            offs = (rng.below(9, 1 << (bd - 5))).astype(np.int16)
            offs[3:5] *= -1
            dst = np.full((size + 1) * S, 6, dtype=np.uint16)
            both(o.sao.band_filter[i], r.sao.band_filter[i], [dst, src],
                 lambda f, d, s: f(p8(d), p8(s), S * 2, S * 2, offs.ctypes.data_as(i16), (i * 7) % 32, w, size))
            for eo in range(4):
                both(o.sao.edge_filter[i], r.sao.edge_filter[i], [dst, src],
                     lambda f, d, s: f(p8(d), p8(s, (S + 16) * 2), S * 2, offs.ctypes.data_as(i16), eo, w, size))


def test_alf_entries(tables):
    bd, o, r = tables
    rng = R(5)
    i16 = C.POINTER(C.c_int16)
    ip = C.POINTER(C.c_int)
    clip_set = np.array([1 << bd, 1 << (bd - 3), 1 << (bd - 5), 1 << (bd - 7)], dtype=np.int16)
    for (w, h) in [(4, 4), (128, 128), (64, 32), (20, 124), (128, 4)]:
        S = 128 + 16
        src = rng.pix((128 + 6) * S, bd)
        off = (3 * S + 3) * 2
        nb = (w // 4) * (h // 4)
        filt = (rng.below(nb * 12, 256) - 128).astype(np.int16)
        clip = clip_set[rng.below(nb * 12, 4)]
        dst = np.full((h + 1) * (S + 4), 6, dtype=np.uint16)
        for vb in (124, h - 4, 60):
            both(o.alf.filter[0], r.alf.filter[0], [dst, src], lambda f, d, s: f(p8(d), (S + 4) * 2, p8(s, off), S * 2, w, h, filt.ctypes.data_as(i16), clip.ctypes.data_as(i16), vb))
            both(o.alf.filter[1], r.alf.filter[1], [dst, src], lambda f, d, s: f(p8(d), (S + 4) * 2, p8(s, off), S * 2, w, h, filt.ctypes.data_as(i16), clip.ctypes.data_as(i16), vb - 2))
            ci, ti = np.full(nb + 1, -1, dtype=np.int32), np.full(nb + 1, -1, dtype=np.int32)
            grad = np.zeros(66 * 66 * 4, dtype=np.int32)
            got = both(o.alf.classify, r.alf.classify, [ci, ti, src],
                       lambda f, a, b, s: f(a.ctypes.data_as(ip), b.ctypes.data_as(ip), p8(s, off), S * 2, w, h, vb, grad.ctypes.data_as(ip)))
        cls, tr = got[0][:nb].copy(), got[1][:nb].copy()
        cset = (rng.below(25 * 12, 257) - 128).astype(np.int16)
        cidx = rng.below(25 * 12, 4).astype(np.uint8)
        cmap = rng.below(25, 25).astype(np.uint8)
        co, cl = np.zeros(nb * 12 + 4, dtype=np.int16), np.zeros(nb * 12 + 4, dtype=np.int16)
        both(o.alf.recon_coeff_and_clip, r.alf.recon_coeff_and_clip, [co, cl],
             lambda f, a, b: f(a.ctypes.data_as(i16), b.ctypes.data_as(i16), cls.ctypes.data_as(ip), tr.ctypes.data_as(ip), nb,
                               cset.ctypes.data_as(i16), cidx.ctypes.data_as(C.POINTER(C.c_uint8)), cmap.ctypes.data_as(C.POINTER(C.c_uint8))))
        # CC-ALF on a chroma block of (w/2, h/2) from the luma around it
        cw, chh = w // 2, h // 2
        cdst = rng.pix((chh + 1) * (cw + 4), bd)
        ccf = np.array([(-1) ** k * (1 << (k % 7)) for k in range(7)], dtype=np.int16)
        for (hs, vs) in [(1, 1), (0, 0)]:
            ww, hh = (cw, chh) if hs else (min(w, 64), min(h, 64))
            cd = rng.pix((hh + 1) * (ww + 4), bd)
            both(o.alf.filter_cc, r.alf.filter_cc, [cd, src], lambda f, d, s: f(p8(d), (ww + 4) * 2, p8(s, off), S * 2, ww, hh, hs, vs, ccf.ctypes.data_as(i16), 124 if vs else 60))


def test_lf_entries(tables):
    bd, o, r = tables
    rng = R(6)
    i32, u8 = C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
    S = 48
    checked = 0
    for it in range(400):
        smooth = it % 3 != 0
        base = rng.pix(1, bd)[0]
        v = (base + rng.below(32 * S, 9 if smooth else 200).astype(np.int64) - 4).reshape(32, S)
        if it % 5 == 0:
            v[:, 24:] += 20
            v[16:, :] += 13
        pic = np.clip(v, 0, (1 << bd) - 1).astype(np.uint16).reshape(-1)
        org = (16 * S + 24) * 2
        qp = 20 + rng.one(40)
        beta = np.array([synth.BETA_TABLE[min(qp, 63)]] * 4, dtype=np.int32)
        tc = np.array([synth.TC_TABLE[min(qp + 2 * rng.one(2), 65)], synth.TC_TABLE[min(qp, 65)] * (it % 7 != 0), 3, 7], dtype=np.int32)
        no_p, no_q = rng.below(4, 5) == 0, rng.below(4, 5) == 0
        no_p, no_q = no_p.astype(np.uint8), no_q.astype(np.uint8)
        lens = np.array([1, 2, 3, 5, 7], dtype=np.uint8)
        mp, mq = lens[rng.below(4, 5)], lens[rng.below(4, 5)]
        hce = rng.one(2)
        for d in range(2):
            try:
                both(o.lf.filter_luma[d], r.lf.filter_luma[d], [pic],
                     lambda f, p: f(p8(p, org), S * 2, beta.ctypes.data_as(i32), tc.ctypes.data_as(i32), no_p.ctypes.data_as(u8), no_q.ctypes.data_as(u8),
                                    mp.ctypes.data_as(u8), mq.ctypes.data_as(u8), hce))
            except AssertionError as e:
                a, b = pic.copy(), pic.copy()
                for f, p in ((o.lf.filter_luma[d], a), (r.lf.filter_luma[d], b)):
                    f(p8(p, org), S * 2, beta.ctypes.data_as(i32), tc.ctypes.data_as(i32), no_p.ctypes.data_as(u8), no_q.ctypes.data_as(u8),
                      mp.ctypes.data_as(u8), mq.ctypes.data_as(u8), hce)
                bad = np.argwhere(a != b).reshape(-1)
                raise AssertionError("it %d dir %d beta %s tc %s no_p %s no_q %s mp %s mq %s hce %d: %s; src/ours/ref rows:\n%s\n%s\n%s" % (
                    it, d, beta[:2], tc[:2], no_p[:2], no_q[:2], mp[:2], mq[:2], hce, [(int(i) // S, int(i) % S, int(a[i]), int(b[i])) for i in bad],
                    pic.reshape(32, S)[8:24, 24:32] if d == 0 else pic.reshape(32, S)[16:24, 16:32],
                    a.reshape(32, S)[8:24, 24:32] if d == 0 else a.reshape(32, S)[16:24, 16:32],
                    b.reshape(32, S)[8:24, 24:32] if d == 0 else b.reshape(32, S)[16:24, 16:32])) from e
            cl = np.array([0, 1, 3], dtype=np.uint8)
            cp, cq = cl[rng.below(4, 3)], cl[rng.below(4, 3)]
            sh = rng.one(2)
            both(o.lf.filter_chroma[d], r.lf.filter_chroma[d], [pic],
                 lambda f, p: f(p8(p, org), S * 2, beta.ctypes.data_as(i32), tc.ctypes.data_as(i32), no_p.ctypes.data_as(u8), no_q.ctypes.data_as(u8),
                                cp.ctypes.data_as(u8), cq.ctypes.data_as(u8), sh))
            both(o.lf.ladf_level[d], r.lf.ladf_level[d], [pic], lambda f, p: f(p8(p, org), S * 2))
            checked += 3
    assert checked == 2400


def test_joint_residuals(tables):
    bd, o, r = tables
    rng = R(7)
    ip = C.POINTER(C.c_int)
    for (w, h) in [(2, 2), (4, 8), (32, 32), (64, 16)]:
        for (sign, shift) in [(1, 0), (-1, 0), (1, 1), (-1, 1)]:
            res = (rng.below(w * h, 8193) - 4096).astype(np.int32)
            pic = rng.pix((h + 1) * (w + 4), bd)
            both(o.itx.add_residual_joint, r.itx.add_residual_joint, [pic, res], lambda f, p, q: f(p8(p), q.ctypes.data_as(ip), w, h, (w + 4) * 2, sign, shift))
            both(o.itx.pred_residual_joint, r.itx.pred_residual_joint, [res], lambda f, q: f(q.ctypes.data_as(ip), w, h, sign, shift))


def test_intra_leaf_predictors(tables):
    """intra.pred_planar / pred_dc / pred_v / pred_h / pred_angular_v / pred_angular_h / pred_mip with the edge pointers
    IntraEdgeParams carries: arrays of 6 * 64 + 5 samples, pointer 67 in (vvcdsp.c:200-207); stride in samples."""
    bd, o, r = tables
    rng = R(11)
    S, OFF = 80, 67
    sizes = [4, 8, 16, 32, 64]
    checked = 0
    for it in range(160):
        w, h = sizes[rng.one(5)], sizes[rng.one(5)]
        top, left = rng.pix(389, bd), rng.pix(389, bd)
        pic = rng.pix((h + 2) * S, bd)
        pt, pl = lambda a: p8(a, OFF * 2), lambda a: p8(a, OFF * 2)
        kind = it % 7
        if kind == 0:
            both(o.intra.pred_planar, r.intra.pred_planar, [pic, top, left], lambda f, p, t, l: f(p8(p, 2 * (S + 4)), pt(t), pl(l), w, h, S))
        elif kind == 1:
            both(o.intra.pred_dc, r.intra.pred_dc, [pic, top, left], lambda f, p, t, l: f(p8(p, 2 * (S + 4)), pt(t), pl(l), w, h, S))
        elif kind == 2:
            both(o.intra.pred_v, r.intra.pred_v, [pic, top], lambda f, p, t: f(p8(p, 2 * (S + 4)), pt(t), w, h, S))
        elif kind == 3:
            both(o.intra.pred_h, r.intra.pred_h, [pic, left], lambda f, p, l: f(p8(p, 2 * (S + 4)), pl(l), w, h, S))
        elif kind == 4:
            size_id = 0 if (w == 4 and h == 4) else (1 if (w == 4 or h == 4 or (w == 8 and h == 8)) else 2)
            mode, tr = rng.one((16, 8, 6)[size_id]), rng.one(2)
            both(o.intra.pred_mip, r.intra.pred_mip, [pic, top, left], lambda f, p, t, l: f(p8(p, 2 * (S + 4)), pt(t), pl(l), w, h, S, mode, tr))
        else:
            vertical = kind == 5
            c_idx = rng.one(3)
            ref_idx = 0 if c_idx else rng.one(3)
            while True:         # a (mode, shape) pair whose reads stay inside the arrays, as every pair the decoder produces does
                mode = (34 + rng.one(47)) if vertical else (-14 + rng.one(48))
                mode = {50: 51, 0: 2, 1: 3, 18: 19}.get(mode, mode)
                along, across = (w, h) if vertical else (h, w)
                reach = (((across + 1 + ref_idx) * abs(synth.intra_angle(mode))) >> 5) + ref_idx + 6
                if along + reach < 190 and reach < 60:
                    break
            flt = rng.one(2)
            pdpc = 0
            if w >= 4 and h >= 4 and not ref_idx and not (18 < mode < 50) and rng.one(4):
                ns, _ = synth.intra_nscale(w, h, mode)
                pdpc = int(ns >= 0)
            fo, fr = (o.intra.pred_angular_v, r.intra.pred_angular_v) if vertical else (o.intra.pred_angular_h, r.intra.pred_angular_h)
            both(fo, fr, [pic, top, left], lambda f, p, t, l: f(p8(p, 2 * (S + 4)), pt(t), pl(l), w, h, S, c_idx, mode, ref_idx, flt, pdpc))
        checked += 1
    assert checked == 160
