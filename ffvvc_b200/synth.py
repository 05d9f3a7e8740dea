"""Seeded synthetic inputs shared by the parity tests and bench.py.

PRNG: the 32-bit LCG s = s*1664525 + 1013904223, value = s >> 8, vectorised with numpy so
CPU oracle, reference and CUDA runs see identical bytes.  Distributions follow the reference's
own checkasm generators where one exists (tests/checkasm/vvc_alf.c:32-79, vvc_sao.c:51-119,
vvc_mc.c:38-327, vvc_itx.c:25-75).
"""
import numpy as np

from . import abi


class LCG:
    A, Cc = 1664525, 1013904223

    def __init__(self, seed=12345):
        self.s = np.uint64(seed & 0xFFFFFFFF)

    def take(self, n):
        """n successive outputs (uint32 array of s >> 8)."""
        n = int(n)
        # closed form jump: s_k = A^k s_0 + C (A^k - 1)/(A - 1)  (mod 2^32), built by doubling
        out = np.empty(n, dtype=np.uint64)
        a_pow = np.ones(n, dtype=np.uint64)
        c_acc = np.zeros(n, dtype=np.uint64)
        k = np.arange(1, n + 1, dtype=np.uint64)
        a, c = np.uint64(self.A), np.uint64(self.Cc)
        mask = np.uint64(0xFFFFFFFF)
        bit = 0
        while (1 << bit) <= n:
            sel = ((k >> np.uint64(bit)) & np.uint64(1)).astype(bool)
            # apply current (a, c) step to the selected lanes: x -> a*x + c
            a_pow[sel] = (a_pow[sel] * a) & mask
            c_acc[sel] = (c_acc[sel] * a + c) & mask
            c = (c * a + c) & mask
            a = (a * a) & mask
            bit += 1
        out = (a_pow * self.s + c_acc) & mask
        self.s = out[-1] if n else self.s
        return (out >> np.uint64(8)).astype(np.uint32)

    def below(self, n, m):
        return (self.take(n) % np.uint32(m)).astype(np.int64)


def uniform_planes(geom, seed=12345):
    """D_uniform: every sample = rnd & ((1 << bit_depth) - 1)  (checkasm-equivalent)."""
    rng = LCG(seed)
    planes = abi.alloc_planes(geom)
    mask = (1 << geom.bit_depth) - 1
    for c, p in enumerate(planes):
        w, h = geom.plane_wh(c)
        v = rng.take(geom.batch * h * w) & np.uint32(mask)
        p[:, :, :w] = v.reshape(geom.batch, h, w).astype(np.uint16)
    return planes


def struct_planes(geom, seed=12345):
    """D_struct: smooth sinusoid mix + 8x8 blockwise DC steps + small noise, clipped to range.

    Needed because on uniform noise deblocking decisions never fire and ALF classes collapse.
    """
    rng = LCG(seed)
    planes = abi.alloc_planes(geom)
    maxv = (1 << geom.bit_depth) - 1
    sc = (1 << geom.bit_depth) / 1024.0
    for c, p in enumerate(planes):
        w, h = geom.plane_wh(c)
        for k in range(geom.batch):
            ph = rng.take(4).astype(np.float64) / float(1 << 24) * 6.28318
            y, x = np.mgrid[0:h, 0:w].astype(np.float64)
            zoom = 1.0 if c == 0 else 2.0
            base = 512 + 300 * np.sin(0.011 * zoom * x + 0.017 * zoom * y + ph[0]) \
                + 120 * np.sin(0.053 * zoom * x - 0.041 * zoom * y + ph[1])
            bh, bw = (h + 7) // 8, (w + 7) // 8
            steps = (rng.below(bh * bw, 49) - 24).reshape(bh, bw)
            dc = np.kron(steps, np.ones((8, 8), dtype=np.int64))[:h, :w]
            noise = (rng.below(h * w, 7) - 3).reshape(h, w)
            v = np.clip(np.floor(base * sc).astype(np.int64) + dc + noise, 0, maxv)
            p[k, :, :w] = v.astype(np.uint16)
    return planes


def alf_params(geom, seed=777, all_on=True, coeffs="int8"):
    """Per-CTB ALF parameters + filter sets, SURVEY.md 8(d) config 1 distributions.

    coeffs: "int8" = checkasm's draw (-128..127, AlfCtbFiltSetIdxY 0..16; the golden fixtures depend on it);
    "full" = the whole legal range -128..+128 (alf_luma_coeff_abs / alf_chroma_coeff_abs 0..128,
    cbs_h266_syntax_template.c:2285,2314; +-abs stored by vvc_ps.c:803-808) in all 8 APS slots, AlfCtbFiltSetIdxY 0..23;
    "max" = every tap +128 or -128."""
    rng = LCG(seed)
    n = geom.ctb_count * geom.batch
    ctbs = np.zeros(n, dtype=abi.ALF_CTB_DTYPE)
    if all_on:
        ctbs["ctb_flag"][:] = 1
    else:
        ctbs["ctb_flag"][:] = rng.below(n * 3, 2).reshape(n, 3)
    ctbs["filt_set_idx_y"] = rng.below(n, 17 if coeffs == "int8" else 24)
    if coeffs != "int8":        # every other CTB takes an APS slot, so that small pictures reach them too
        ctbs["filt_set_idx_y"][::2] = 16 + ctbs["filt_set_idx_y"][::2] % 8
    ctbs["chroma_alt_idx"][:] = rng.below(n * 2, 8).reshape(n, 2)
    ctbs["cc_idc"][:] = rng.below(n * 2, 5).reshape(n, 2)
    sets = np.zeros(1, dtype=abi.ALF_SETS_DTYPE)
    if coeffs == "int8":
        # APS luma: random int8-range coefficients (tests/checkasm/vvc_alf.c:55-59), clip idx rnd%4
        sets["luma_coeff"][0] = (rng.below(8 * 25 * 12, 256) - 128).reshape(8, 25, 12)
        sets["luma_clip_idx"][0] = rng.below(8 * 25 * 12, 4).reshape(8, 25, 12)
        sets["chroma_coeff"][0] = (rng.below(8 * 6, 256) - 128).reshape(8, 6)
        sets["chroma_clip_idx"][0] = rng.below(8 * 6, 4).reshape(8, 6)
    else:
        if coeffs == "full":
            sets["luma_coeff"][0] = (rng.below(8 * 25 * 12, 257) - 128).reshape(8, 25, 12)
            sets["chroma_coeff"][0] = (rng.below(8 * 6, 257) - 128).reshape(8, 6)
            # a quarter of the filters get at least one +128 so the wide path is always exercised
            hit = rng.below(8 * 25, 4).reshape(8, 25) == 0
            tap = rng.below(8 * 25, 12).reshape(8, 25)
            for a, b in zip(*np.nonzero(hit)):
                sets["luma_coeff"][0][a, b, tap[a, b]] = 128
            sets["chroma_coeff"][0][::2, rng.below(1, 6)[0]] = 128
        else:
            assert coeffs == "max"
            sets["luma_coeff"][0] = (rng.below(8 * 25 * 12, 2) * 256 - 128).reshape(8, 25, 12)
            sets["chroma_coeff"][0] = (rng.below(8 * 6, 2) * 256 - 128).reshape(8, 6)
        sets["luma_clip_idx"][0] = rng.below(8 * 25 * 12, 4).reshape(8, 25, 12)
        sets["chroma_clip_idx"][0] = rng.below(8 * 6, 4).reshape(8, 6)
    # CC-ALF coefficients are 0 or +-2^k, k <= 6 (vvc_ps.c:810-819)
    mag = rng.below(2 * 5 * 7, 8)
    sgn = rng.below(2 * 5 * 7, 2) * 2 - 1
    cc = np.where(mag == 0, 0, sgn * (1 << np.maximum(mag - 1, 0)))
    sets["cc_coeff"][0] = cc.reshape(2, 5, 7)
    return ctbs, sets


# ---------------------------------------------------------------------------------------------
# Deblocking inputs: a random transform-block partition, and from it the per-segment
# (tc, beta, max lengths) the reference driver derives (vvc_filter.c:374-397, 795-826, 861-1003).
# ---------------------------------------------------------------------------------------------
TC_TABLE = np.array([0] * 18 + [3, 4, 4, 4, 4, 5, 5, 5, 5, 7, 7, 8, 9, 10, 10, 11, 13, 14, 15, 17, 19, 21, 24, 25, 29, 33,
                                36, 41, 45, 51, 57, 64, 71, 80, 89, 100, 112, 125, 141, 157, 177, 198, 222, 250, 280, 314,
                                352, 395], dtype=np.int64)          # Table 43, vvc_filter.c:38-44
BETA_TABLE = np.array([0] * 16 + list(range(6, 19)) + list(range(20, 90, 2)), dtype=np.int64)   # vvc_filter.c:47-52
assert len(TC_TABLE) == 66 and len(BETA_TABLE) == 64


def tb_partition(geom, rng, stop_p=0.35):
    """Random aligned TB partition per 4x4 luma unit: returns (log2 tb width, log2 tb height).

    Blocks start at 64x64 and are halved alternately in width and height (order chosen per
    64x64 block) until a random stop, so sizes 4..64 and 1:2 / 2:1 shapes occur.
    """
    uh, uw = (geom.height + 3) // 4, (geom.width + 3) // 4
    uy, ux = np.mgrid[0:uh, 0:uw]
    lw = np.full((uh, uw), 6, dtype=np.int64)
    lh = np.full((uh, uw), 6, dtype=np.int64)
    alive = np.ones((uh, uw), dtype=bool)
    g64h, g64w = (uh + 15) // 16, (uw + 15) // 16
    width_first = rng.below(g64h * g64w, 2).reshape(g64h, g64w)[uy >> 4, ux >> 4].astype(bool)
    for stage in range(8):
        # block id of every unit at its current shape
        by, bx = uy >> (lh - 2), ux >> (lw - 2)
        gh, gw = int(by.max()) + 1, int(bx.max()) + 1
        dec = (rng.below(gh * gw, 1000) >= int(stop_p * 1000)).reshape(gh, gw)
        split = alive & dec[by, bx]
        halve_w = (stage % 2 == 0) == width_first
        lw = np.where(split & halve_w, lw - 1, lw)
        lh = np.where(split & ~halve_w, lh - 1, lh)
        alive = split
    return lw, lh


def _edge_records(bs, qp_edge, len_p, len_q):
    rec = np.zeros(bs.shape, dtype=abi.DBK_EDGE_DTYPE)
    tc = TC_TABLE[np.clip(qp_edge + 2 * (bs - 1), 0, 65)]
    rec["tc"] = np.where(bs > 0, tc, 0)
    rec["beta"] = np.where(bs > 0, BETA_TABLE[np.clip(qp_edge, 0, 63)], 0)
    rec["max_len"] = np.where(bs > 0, len_p | (len_q << 4), 0)
    return rec


def deblock_maps(geom, seed=4242, qp_base=22, qp_span=21):
    """Edge maps [dir][c] (batch, rows, pitch) following SURVEY.md 8(d) config 2:
    bs in {0,1,2} with p = {.45,.35,.20} on the 8x8 grid and {.85,.10,.05} on off-grid 4-lines,
    qp = qp_base + rnd % qp_span per 16x16, luma max lengths from the TB sizes
    (derive_max_filter_length_luma), chroma lengths from chroma TB sizes and bs."""
    rng = LCG(seed)
    uh, uw = (geom.height + 3) // 4, (geom.width + 3) // 4
    out = [[None] * 3, [None] * 3]
    for d in range(2):
        for c in range(3):
            rows, pitch = abi.deblock_map_shape(geom, d, c)
            out[d][c] = np.zeros((geom.batch, rows, pitch), dtype=abi.DBK_EDGE_DTYPE)
    ctb4 = geom.ctb_size // 4
    for k in range(geom.batch):
        lw, lh = tb_partition(geom, rng)
        uy, ux = np.mgrid[0:uh, 0:uw]
        qp16 = qp_base + rng.below(((uh + 3) // 4) * ((uw + 3) // 4), qp_span).reshape((uh + 3) // 4, (uw + 3) // 4)
        qp = qp16[uy >> 2, ux >> 2]
        for d in (1, 0):
            size = (1 << lw) if d else (1 << lh)          # TB size across the edge
            pos = ux if d else uy
            is_edge = ((pos * 4) % size == 0) & (pos > 0)
            size_q = size
            size_p = np.roll(size, 1, axis=1 if d else 0)
            qp_p = np.roll(qp, 1, axis=1 if d else 0)
            on8 = (pos % 2 == 0)
            r = rng.below(uh * uw, 100).reshape(uh, uw)
            bs = np.where(on8, (r >= 45).astype(np.int64) + (r >= 80), (r >= 85).astype(np.int64) + (r >= 95))
            bs = np.where(is_edge, bs, 0)
            # ---- luma (vvc_filter.c:374-397) ----
            small = (size_p <= 4) | (size_q <= 4)
            lp = np.where(small, 1, np.where(size_p >= 32, 7, 3))
            lq = np.where(small, 1, np.where(size_q >= 32, 7, 3))
            # sub-block (affine/sbTMVP) CUs cap the q side, merge-subblock/intra-affine neighbours the p side (:392-396)
            cap = rng.below(uh * uw, 10).reshape(uh, uw)
            lq = np.where(cap == 0, np.minimum(lq, 5), lq)
            lp = np.where(cap == 1, np.minimum(lp, 5), lp)
            qe = (qp + qp_p + 1) >> 1
            rec = _edge_records(bs, qe, lp, lq)
            # luma map rows/cols are exactly the 4x4 units
            out[d][0][k] = rec
            # ---- chroma 4:2:0: edges on the 8-sample chroma grid = 16 luma, i.e. every 4th unit ----
            if geom.chroma_format_idc:
                csize_p, csize_q = size_p >> 1, size_q >> 1
                ctu_edge = (uy % ctb4 == 0) if d == 0 else np.zeros_like(on8)
                big = (csize_p >= 8) & (csize_q >= 8)
                clp = np.where(big, np.where(ctu_edge, 1, 3), (bs == 2).astype(np.int64))
                clq = np.where(big, 3, (bs == 2).astype(np.int64))
                for c in (1, 2):
                    cqe = qe + (1 if c == 1 else -1)     # distinct Cb/Cr qp, as separate qp tables give
                    crec = _edge_records(bs, cqe, clp, clq)
                    rows, pitch = abi.deblock_map_shape(geom, d, c)
                    if d:   # vertical: x every 16 luma = 4 units; segment = 2 chroma lines = 1 unit row
                        out[d][c][k] = crec[:rows, ::4][:, :pitch]
                    else:   # horizontal: y every 4 units; segment = 2 chroma columns = 1 unit column
                        out[d][c][k] = crec[::4, :pitch][:rows]
    return out


def deblock_side_info(geom, seed=515, ladf=True):
    """What the parser knows about a picture ring, as the list inputs of vvc_cuda_deblock_params_frame: coding units on a
    random partition (intra / inter / CIIP / sub-block motion), each cut into one, two or four transform units with random
    coded flags, joint CbCr, BDPCM, per-unit QPs; motion rectangles with vectors from a small palette plus jitter so that
    neighbours differ by less and by more than half a sample, different reference pictures, uni / bi; per-CTB beta / tc
    offsets and unfilterable tile borders; LADF intervals.  Returns (tus, mvfs, ctbs, VVCCudaDbkParams)."""
    rng = LCG(seed)
    tus, mvfs = [], []
    for k in range(geom.batch):
        lw, lh = tb_partition(geom, rng, stop_p=0.4)
        uh, uw = lw.shape
        uy, ux = np.mgrid[0:uh, 0:uw]
        origin = ((ux * 4) % (1 << lw) == 0) & ((uy * 4) % (1 << lh) == 0)
        x0, y0, l2w, l2h = ux[origin] * 4, uy[origin] * 4, lw[origin], lh[origin]
        inside = (x0 + (1 << l2w) <= geom.width) & (y0 + (1 << l2h) <= geom.height)
        # blocks that stick out of the picture are re-cut into 4x4 units (every unit of the picture belongs to a block)
        out = ~inside
        ex, ey = [], []
        for bx, by, bw, bh in zip(x0[out], y0[out], l2w[out], l2h[out]):
            for yy in range(by, min(by + (1 << bh), geom.height), 4):
                for xx in range(bx, min(bx + (1 << bw), geom.width), 4):
                    ex.append(xx); ey.append(yy)
        x0 = np.concatenate([x0[inside], np.array(ex, dtype=np.int64)])
        y0 = np.concatenate([y0[inside], np.array(ey, dtype=np.int64)])
        l2w = np.concatenate([l2w[inside], np.full(len(ex), 2, dtype=np.int64)])
        l2h = np.concatenate([l2h[inside], np.full(len(ex), 2, dtype=np.int64)])
        # a coding block never exceeds the CTB
        while (l2w > geom.ctb_log2).any() or (l2h > geom.ctb_log2).any():
            big_w = l2w > geom.ctb_log2
            x0 = np.concatenate([x0, x0[big_w] + (1 << (l2w[big_w] - 1))]); y0 = np.concatenate([y0, y0[big_w]])
            l2h = np.concatenate([l2h, l2h[big_w]]); l2w = np.concatenate([np.where(big_w, l2w - 1, l2w), l2w[big_w] - 1])
            big_h = l2h > geom.ctb_log2
            x0 = np.concatenate([x0, x0[big_h]]); y0 = np.concatenate([y0, y0[big_h] + (1 << (l2h[big_h] - 1))])
            l2w = np.concatenate([l2w, l2w[big_h]]); l2h = np.concatenate([np.where(big_h, l2h - 1, l2h), l2h[big_h] - 1])
        n = len(x0)
        R = lambda m: rng.below(n, m)
        kind, split, sb_r, pal, jit, prd, refa, refb, qpd, bdp = R(100), R(8), R(100), R(6), R(1 << 16), R(3), R(3), R(3), R(7), R(20)
        cbf = rng.below(4 * n, 8).reshape(n, 4)
        palette = np.array([[0, 0], [37, -12], [-150, 64], [8, 8], [1024, -700], [-5, 3]])
        for i in range(n):
            x, y, w, h = int(x0[i]), int(y0[i]), 1 << int(l2w[i]), 1 << int(l2h[i])
            intra = kind[i] < 20
            ciip = (not intra) and kind[i] < 26 and w * h >= 64
            sb = (not intra) and (not ciip) and sb_r[i] < 20 and w >= 8 and h >= 8
            qp_y = 27 + int(qpd[i]) - 3
            # transform units: the whole block, two halves or four quarters
            cuts = [(0, 0, w, h)]
            if split[i] == 0 and w >= 8:
                cuts = [(0, 0, w // 2, h), (w // 2, 0, w // 2, h)]
            elif split[i] == 1 and h >= 8:
                cuts = [(0, 0, w, h // 2), (0, h // 2, w, h // 2)]
            elif split[i] == 2 and w >= 8 and h >= 8:
                cuts = [(0, 0, w // 2, h // 2), (w // 2, 0, w // 2, h // 2), (0, h // 2, w // 2, h // 2), (w // 2, h // 2, w // 2, h // 2)]
            elif split[i] == 3 and w >= 16:        # a quarter and three quarters (SBT-like), offsets that are not multiples of 8
                cuts = [(0, 0, w // 4, h), (w // 4, 0, 3 * w // 4, h)] if w // 4 >= 4 else cuts
            for j, (dx, dy, tw, th) in enumerate(cuts):
                parts = [(dx, dy, tw, th)]
                if tw & (tw - 1):                   # 3/4 of a power of two: two records of power-of-two width
                    parts = [(dx, dy, tw * 2 // 3, th), (dx + tw * 2 // 3, dy, tw // 3, th)]
                for (px, py, pw, ph) in parts:
                    t = np.zeros(1, dtype=abi.DBK_TU_DTYPE)
                    t["x0"], t["y0"], t["log2_w"], t["log2_h"], t["pic"] = x + px, y + py, pw.bit_length() - 1, ph.bit_length() - 1, k
                    t["planes"] = abi.DBK_TU_LUMA | (abi.DBK_TU_CHROMA if geom.chroma_format_idc else 0)
                    f = int(cbf[i, j % 4])
                    fl = (abi.DBK_CBF_Y if f & 1 else 0) | (abi.DBK_CBF_CB if f & 2 else 0) | (abi.DBK_CBF_CR if f & 4 else 0)
                    if f == 6 and bdp[i] < 4:
                        fl |= abi.DBK_JOINT
                    if intra and bdp[i] == 0:
                        fl |= abi.DBK_BDPCM_Y | abi.DBK_BDPCM_C
                    t["flags"] = fl
                    t["qp"][0] = (qp_y, qp_y + 12 + (j % 3) - 1, qp_y + 12 - (j % 2))
                    t["cu_flags"] = abi.DBK_CU_SUBBLOCK if sb else 0
                    t["cu_dx"], t["cu_dy"], t["cb_log2_w"], t["cb_log2_h"] = px // 4, py // 4, int(l2w[i]), int(l2h[i])
                    tus.append(t)
            # motion
            base = palette[int(pal[i])]
            pf = 0 if intra else int(prd[i]) + 1
            def rec(rx, ry, rw, rh, mv0, mv1):
                m = np.zeros(1, dtype=abi.DBK_MVF_DTYPE)
                m["x0"], m["y0"], m["w4"], m["h4"], m["pred_flag"], m["ciip_flag"], m["pic"] = rx, ry, rw // 4, rh // 4, pf, int(ciip), k
                m["ref_pic"][0] = (int(refa[i]), int(refb[i]))
                m["mv"][0] = (mv0, mv1)
                mvfs.append(m)
            j0 = np.array([int(jit[i]) % 13 - 6, (int(jit[i]) >> 4) % 13 - 6])
            if sb:
                for sy in range(0, h, 4):
                    for sx in range(0, w, 4):
                        d = np.array([5 * (sx // 4), -5 * (sy // 4)])
                        rec(x + sx, y + sy, 4, 4, base + j0 + d, base - j0 + 2 * d)
            else:
                rec(x, y, w, h, base + j0, base - j0 if pal[i] & 1 else base + j0)
    ctbs = np.zeros(geom.ctb_count * geom.batch, dtype=abi.DBK_CTB_DTYPE)
    nc = len(ctbs)
    ctbs["beta_offset"] = (rng.below(nc * 3, 25) - 12).reshape(nc, 3)
    ctbs["tc_offset"] = (rng.below(nc * 3, 25) - 12).reshape(nc, 3)
    ctbs["no_left"] = rng.below(nc, 8) == 0
    ctbs["no_top"] = rng.below(nc, 8) == 0
    p = abi.VVCCudaDbkParams()
    p.qp_bd_offset = 6 * (geom.bit_depth - 8)
    p.ladf_enabled = int(ladf)
    p.num_ladf_intervals = 4
    p.ladf_lowest_interval_qp_offset = -2
    bounds = (0, 1 << (geom.bit_depth - 2), 1 << (geom.bit_depth - 1), 3 << (geom.bit_depth - 2), 0)
    for i in range(4):
        p.ladf_qp_offset[i] = (3, -1, 2, 0)[i]
    for i in range(5):
        p.ladf_interval_lower_bound[i] = bounds[i]
    return np.concatenate(tus), np.concatenate(mvfs), ctbs, p


def sao_params(geom, seed=99, with_restore=False):
    """Per-CTB SAO parameters, SURVEY.md 8(d) config 2: type rnd%3, band position rnd%32,
    eo rnd%4, offsets rnd % (1 << (bd-5)) with the edge-class signs (+,+,-,-) the parser applies
    (vvc_ctu.c:2202-2211)."""
    rng = LCG(seed)
    n = geom.ctb_count * geom.batch
    p = np.zeros(n, dtype=abi.SAO_CTB_DTYPE)
    p["type_idx"][:] = rng.below(n * 3, 3).reshape(n, 3)
    p["band_position"][:] = rng.below(n * 3, 32).reshape(n, 3)
    p["eo_class"][:] = rng.below(n * 3, 4).reshape(n, 3)
    mag = rng.below(n * 3 * 4, 1 << (geom.bit_depth - 5)).reshape(n, 3, 4)
    sign_band = rng.below(n * 3 * 4, 2).reshape(n, 3, 4) * 2 - 1
    edge_sign = np.array([1, 1, -1, -1])
    is_edge = (p["type_idx"] == 2)[:, :, None]
    p["offset_val"][:, :, 1:] = np.where(is_edge, mag * edge_sign, mag * sign_band)
    if with_restore:
        p["restore"] = 1
        flags = rng.below(n, 256).astype(np.uint8)
        cols, rows = geom.ctb_cols, geom.ctb_rows
        idx = np.arange(n) % (cols * rows)
        cx, cy = idx % cols, idx // cols
        # flags are only ever set towards neighbours that exist (vvc_filter.c:181-212)
        left, top, right, bottom = cx == 0, cy == 0, cx == cols - 1, cy == rows - 1
        clear = np.zeros(n, dtype=np.uint8)
        clear |= np.where(left, 0x01 | 0x10 | 0x80, 0).astype(np.uint8)
        clear |= np.where(right, 0x02 | 0x20 | 0x40, 0).astype(np.uint8)
        clear |= np.where(top, 0x04 | 0x10 | 0x20, 0).astype(np.uint8)
        clear |= np.where(bottom, 0x08 | 0x40 | 0x80, 0).astype(np.uint8)
        p["no_filter"] = flags & ~clear
    return p


# ---------------------------------------------------------------------------------------------
# Residual stage inputs (coefficient/nz distributions of
# tests/checkasm/vvc_itx.c:25-75)
# ---------------------------------------------------------------------------------------------
def tb_list(geom, seed=31337, lfnst_set_of=None, extras=True, saturate=True, square=False, min_log2=2):
    """Tile every picture of the ring with transform blocks (luma + both chroma planes).

    Returns (tbs, coeffs): TB_DTYPE records and the dense int32 coefficient buffer.
    lfnst_set_of: callable predModeIntra -> transform set (the oracle's table accessor); when None no
    TB uses LFNST.  extras: append 1-D and STORE_RESIDUAL blocks (do not touch the picture).
    """
    rng = LCG(seed)
    recs = []
    for k in range(geom.batch):
        lw, lh = tb_partition(geom, rng, stop_p=0.45)
        if square:                       # a quadtree: every block cut into squares of its shorter side
            lw = lh = np.maximum(np.minimum(lw, lh), min_log2)
        uh, uw = lw.shape
        uy, ux = np.mgrid[0:uh, 0:uw]
        origin = ((ux * 4) % (1 << lw) == 0) & ((uy * 4) % (1 << lh) == 0)
        # blocks must lie inside the picture (partition blocks are aligned, pictures are multiples of 8:
        # clip oversize blocks by re-splitting them into 8x8)
        x0, y0 = ux[origin] * 4, uy[origin] * 4
        l2w, l2h = lw[origin], lh[origin]
        inside = (x0 + (1 << l2w) <= geom.width) & (y0 + (1 << l2h) <= geom.height)
        x0, y0, l2w, l2h = x0[inside], y0[inside], l2w[inside], l2h[inside]
        n = len(x0)
        for c in range(3 if geom.chroma_format_idc else 1):
            r = np.zeros(n, dtype=abi.TB_DTYPE)
            sh = 1 if c else 0
            r["x0"], r["y0"] = x0 >> sh, y0 >> sh
            r["log2_w"], r["log2_h"] = l2w - sh, l2h - sh
            r["c_idx"], r["pic"] = c, k
            recs.append(r)
    tbs = np.concatenate(recs)
    n = len(tbs)
    w, h = 1 << tbs["log2_w"].astype(np.int64), 1 << tbs["log2_h"].astype(np.int64)
    luma = tbs["c_idx"] == 0
    # transform types: luma TBs <= 32 draw from {DCT2, DST7, DCT8}^2, others DCT2
    mts_ok_w = luma & (w >= 4) & (w <= 32) & (h <= 32) & (h >= 4)
    trh = np.where(mts_ok_w, rng.below(n, 3), 0)
    trv = np.where(mts_ok_w, rng.below(n, 3), 0)
    tbs["trh"], tbs["trv"] = trh, trv
    tbs["nzw"] = rng.below(n, 1 << 16) % np.minimum(np.where(trh == 0, 32, 16), w) + 1
    tbs["nzh"] = rng.below(n, 1 << 16) % np.minimum(np.where(trv == 0, 32, 16), h) + 1
    # transform skip / BDPCM on 3 % of the blocks up to 32x32
    r = rng.below(n, 100)
    ts = (r < 3) & (w <= 32) & (h <= 32) & (w >= 4) & (h >= 4)
    bd = rng.below(n, 4)
    tbs["flags"] = np.where(ts, abi.TB_TS | np.where(bd == 0, abi.TB_BDPCM, np.where(bd == 1, abi.TB_BDPCM_VERT, 0)), 0)
    tbs["trh"] = np.where(ts, 0, tbs["trh"])
    tbs["trv"] = np.where(ts, 0, tbs["trv"])
    # LFNST on 10 % of the non-skipped blocks >= 4x4 (forces DCT2 x DCT2)
    if lfnst_set_of is not None:
        r = rng.below(n, 100)
        lf = (r >= 90) & ~ts & (w >= 4) & (h >= 4)
        mode = rng.below(n, 95 + 14) - 14                      # predModeIntra incl. wide angles -14..-1, 67..80
        sets = np.array([lfnst_set_of(int(m)) for m in range(-14, 95)])[mode + 14]
        idx = rng.below(n, 2) + 1
        small_in = ((w == 8) & (h == 8)) | ((w == 4) & (h == 4))
        code = idx | (sets << 2) | ((mode > 34).astype(np.int64) << 4) | (small_in.astype(np.int64) << 5)
        tbs["lfnst"] = np.where(lf, code, 0)
        tbs["trh"] = np.where(lf, 0, tbs["trh"])
        tbs["trv"] = np.where(lf, 0, tbs["trv"])
    # joint CbCr on 10 % of the Cb blocks: the co-located Cr block is dropped
    if geom.chroma_format_idc:
        cb = np.nonzero(tbs["c_idx"] == 1)[0]
        cr = np.nonzero(tbs["c_idx"] == 2)[0]
        assert len(cb) == len(cr)
        j = rng.below(len(cb), 10) == 0
        tbs["flags"][cb[j]] |= abi.TB_JOINT
        tbs["joint_c_idx"][cb[j]] = 2
        tbs["joint_sign"][cb[j]] = rng.below(int(j.sum()), 2) * 2 - 1
        tbs["joint_shift"][cb[j]] = rng.below(int(j.sum()), 2)
        keep = np.ones(n, dtype=bool)
        keep[cr[j]] = False
        tbs = tbs[keep]
    if extras:
        # 1-D cells (16, 32, 64 long) and blocks that return their residual instead of adding it
        ex = np.zeros(24, dtype=abi.TB_DTYPE)
        for i in range(24):
            size = (4, 5, 6)[i % 3]
            horiz = (i // 3) % 2
            tr = 0 if size == 6 else (i // 6) % 3
            ex[i]["log2_w"], ex[i]["log2_h"] = (size, 0) if horiz else (0, size)
            ex[i]["trh"], ex[i]["trv"] = (tr, 0) if horiz else (0, tr)
            lim = min(32 if tr == 0 else 16, 1 << size)
            ex[i]["nzw"] = (int(rng.below(1, lim)[0]) + 1) if horiz else 1
            ex[i]["nzh"] = 1 if horiz else (int(rng.below(1, lim)[0]) + 1)
            ex[i]["flags"] = abi.TB_STORE_RESIDUAL
        pick = rng.below(64, len(tbs))
        more = tbs[pick].copy()
        more["flags"] = (more["flags"] & (255 - abi.TB_JOINT)) | abi.TB_STORE_RESIDUAL
        tbs = np.concatenate([tbs, ex, more])
    n = len(tbs)
    w, h = 1 << tbs["log2_w"].astype(np.int64), 1 << tbs["log2_h"].astype(np.int64)
    area = w * h
    off = np.concatenate([[0], np.cumsum(area)])
    tbs["coeff_offset"] = off[:-1]
    total = int(off[-1])
    # coefficients: clip_intp2(rnd, 15) inside the nz window, zero outside (decoder invariant, vvc_cabac.c:2392)
    tb_of = np.repeat(np.arange(n), area)
    local = np.arange(total) - off[tb_of]
    cx, cy = local % w[tb_of], local // w[tb_of]
    full = ((tbs["flags"] & abi.TB_TS) != 0) | (tbs["lfnst"] != 0)
    lf_in = np.where((tbs["lfnst"] >> 5) & 1, 8, 16)
    inside = (cx < tbs["nzw"][tb_of]) & (cy < tbs["nzh"][tb_of])
    inside = np.where(full[tb_of], True, inside)
    raw = rng.take(total).astype(np.int64)
    wide = ((raw << 8) ^ (raw >> 3)) & 0xFFFFFFFF
    wide = np.where(wide >= 1 << 31, wide - (1 << 32), wide)
    sat = np.clip(wide, -32768, 32767)                       # checkasm style: mostly saturated
    mild = (raw % 1025) - 512
    use_sat = (rng.below(n, 2) == 0)[tb_of] if saturate else np.zeros(total, dtype=bool)
    val = np.where(use_sat, sat, mild)
    # LFNST blocks carry only their first 8/16 diagonal-scan coefficients
    diag_x = np.array([0, 0, 1, 0, 1, 2, 0, 1, 2, 3, 1, 2, 3, 2, 3, 3])
    diag_y = np.array([0, 1, 0, 2, 1, 0, 3, 2, 1, 0, 3, 2, 1, 3, 2, 3])
    rank = np.full((4, 4), 99)
    rank[diag_y, diag_x] = np.arange(16)
    in_44 = (cx < 4) & (cy < 4)
    lf_rank = np.where(in_44, rank[np.minimum(cy, 3), np.minimum(cx, 3)], 99)
    is_lf = (tbs["lfnst"] != 0)[tb_of]
    inside = np.where(is_lf, lf_rank < lf_in[tb_of], inside)
    coeffs = np.where(inside, val, 0).astype(np.int32)
    return tbs, coeffs


# Table 38 of H.266 (scaling matrix id by predMode, cIdx, max(nTbW, nTbH)); rows as in derive_scale_m,
# libavcodec/vvc/vvc_intra.c:344-355
SL_IDS = np.array([[[0, 2, 8, 14, 20, 26], [0, 3, 9, 15, 21, 21], [0, 4, 10, 16, 22, 22]],
                   [[0, 5, 11, 17, 23, 27], [0, 6, 12, 18, 24, 24], [1, 7, 13, 19, 25, 25]]])


def tb_for_window(tbs):
    """Make nzw/nzh of a TB list describe the window the residual coder wrote (what the WINDOW16 layout stores):
    transform-skip blocks cover the block, LFNST blocks at least their 4x4 corner."""
    t = tbs.copy()
    w, h = 1 << t["log2_w"].astype(np.int64), 1 << t["log2_h"].astype(np.int64)
    ts = (t["flags"] & abi.TB_TS) != 0
    lf = t["lfnst"] != 0
    t["nzw"] = np.where(ts, w, np.where(lf, np.minimum(w, 4), np.minimum(t["nzw"], w)))
    t["nzh"] = np.where(ts, h, np.where(lf, np.minimum(h, 4), np.minimum(t["nzh"], h)))
    return t


def tb_quant(tbs, seed=77, scaling=True, qp_lo=4, qp_hi=74):
    """Per-TB quantisation records (TB_QUANT_DTYPE) and one scaling list (SCALING_LIST_DTYPE, or None):
    qp uniform in [qp_lo, qp_hi], dependent quantisation on half of the pictures' TBs, explicit scaling
    matrices (Table 38 id by a per-TB intra/inter draw) on 60 % of the non-skipped blocks."""
    rng = LCG(seed)
    n = len(tbs)
    q = np.zeros(n, dtype=abi.TB_QUANT_DTYPE)
    q["qp"] = qp_lo + rng.below(n, qp_hi - qp_lo + 1)
    q["dep_quant"] = rng.below(n, 2)
    sl = None
    if scaling:
        sl = np.zeros(1, dtype=abi.SCALING_LIST_DTYPE)
        sl["matrix_rec"][0] = (1 + rng.below(28 * 64, 255)).reshape(28, 64)
        sl["dc_rec"][0] = 1 + rng.below(14, 255)
        inter = rng.below(n, 2)
        size_idx = np.maximum(tbs["log2_w"], tbs["log2_h"]).astype(np.int64) - 1
        ids = SL_IDS[inter, tbs["c_idx"].astype(np.int64), np.maximum(size_idx, 0)]
        ts = (tbs["flags"] & abi.TB_TS) != 0
        use = (rng.below(n, 100) < 60) & ~ts & (size_idx >= 0)
        q["sl_id"] = np.where(use, ids + 1, 0)
    return q, sl


def lmcs_luts(bit_depth=10, seed=5):
    """A forward/inverse LMCS look-up pair shaped like the ones vvc_ps.c:592-672 derives: 16 bins,
    piecewise linear, monotone; the inverse LUT is the numeric inverse of the forward one."""
    rng = LCG(seed)
    n = 1 << bit_depth
    org = n // 16
    cw = org // 2 + rng.below(16, org)                  # bin code words, each in [org/2, 3*org/2)
    cw[0] = cw[15] = 0                                    # lmcs_min_bin_idx = 1, max = 14
    cw = (cw * (n - 1) // max(int(cw.sum()), 1)).astype(np.int64)
    pivot = np.concatenate([[0], np.cumsum(cw)])
    x = np.arange(n)
    b = np.minimum(x // org, 15)
    fwd = np.clip(pivot[b] + ((x - b * org) * cw[b] + org // 2) // org, 0, n - 1)
    inv = np.clip(np.searchsorted(fwd, x, side="left"), 0, n - 1)
    return fwd.astype(np.uint16), inv.astype(np.uint16)


def lmcs_chroma(geom, tbs, seed=9, literal=False):
    """LMCS chroma residual scaling inputs for a TB list (ph_chroma_residual_scale_flag pictures).

    Returns (tbs2, order, vpdus, params, n_luma): the TB records reordered luma first (order = the permutation applied,
    for per-TB side arrays), every 64x64 (min(CtbSizeY, 64)) VPDU of every picture with the availability of its left / top
    neighbours (0 at the picture border, and at a random quarter of the CTB borders - a slice or tile boundary), VVCLMCS's
    pivots / chroma_scale_coeff / bin range (shaped like vvc_ps.c derives them), and the number of luma blocks.  Chroma
    blocks of more than 4 samples carry 1 + the index of the VPDU of their origin (literal=True: a scale itself)."""
    rng = LCG(seed)
    size = min(geom.ctb_size, 64)
    cols, rows = (geom.width + size - 1) // size, (geom.height + size - 1) // size
    n = cols * rows * geom.batch
    v = np.zeros(n, dtype=abi.LMCS_VPDU_DTYPE)
    idx = np.arange(n)
    v["pic"] = idx // (cols * rows)
    v["x"] = (idx % cols) * size
    v["y"] = ((idx // cols) % rows) * size
    cut = rng.below(2 * n, 4) == 0
    ctb = geom.ctb_size
    v["avail_l"] = (v["x"] > 0) & ~((v["x"] % ctb == 0) & cut[:n])
    v["avail_t"] = (v["y"] > 0) & ~((v["y"] % ctb == 0) & cut[n:])
    bd = geom.bit_depth
    org = (1 << bd) // 16
    p = np.zeros(1, dtype=abi.LMCS_PARAMS_DTYPE)
    cw = org // 2 + rng.below(16, org)
    cw[0] = cw[15] = 0
    p["min_bin_idx"], p["max_bin_idx"] = 1, 14
    cw = (cw * ((1 << bd) - 1) // max(int(cw.sum()), 1)).astype(np.int64)
    p["pivot"][0] = np.concatenate([[0], np.cumsum(cw)])
    p["chroma_scale_coeff"][0] = np.where(cw > 0, (org << 11) // np.maximum(cw + rng.below(16, 15).astype(np.int64) - 7, 1), 1 << 11)
    order = np.argsort(tbs["c_idx"] > 0, kind="stable")
    t2 = tbs[order].copy()
    n_luma = int((t2["c_idx"] == 0).sum())
    ch = (t2["c_idx"] > 0) & ((t2["log2_w"].astype(np.int64) + t2["log2_h"]) > 2)
    sh_x, sh_y = geom.hshift, geom.vshift
    vi = (t2["pic"].astype(np.int64) * rows + ((t2["y0"].astype(np.int64) << sh_y) // size)) * cols + ((t2["x0"].astype(np.int64) << sh_x) // size)
    if literal:
        lit = 256 + rng.below(len(t2), 16128)
        lit[rng.below(len(t2), 16) == 0] = 65535
        t2["chroma_scale"] = np.where(ch, lit, 0)
    else:
        assert n < 65535
        t2["chroma_scale"] = np.where(ch, vi + 1, 0)
    return t2, order, v, p, n_luma


# ---------------------------------------------------------------------------------------------
# Inter prediction inputs: a random-access-like motion field
# ---------------------------------------------------------------------------------------------
_TABLES = {}


def lfnst_set_of(pred_mode_intra):
    """LFNST transform set of an intra mode (wide angles < 0 -> set 1, vvc_itx_1d.c:711), read from the generated
    tables: the callable tb_list() wants, without going through any library."""
    return 1 if pred_mode_intra < 0 else int(spec_table("lfnst_tr_set_index")[pred_mode_intra])


def spec_table(name):
    """One table of ffvvc_b200/csrc/vvc_tables.inc (generated H.266 constants) as a numpy array."""
    if not _TABLES:
        import os
        import re
        src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "vvc_tables.inc")).read()
        for m in re.finditer(r"VVCT_TABLE\((\w+), vvct_(\w+), ((?:\[\d+\])+)\) = \{([^}]*)\}", src):
            dims = [int(d) for d in re.findall(r"\[(\d+)\]", m.group(3))]
            vals = np.array([int(v) for v in m.group(4).replace("\n", " ").split(",") if v.strip()], dtype=np.int64)
            _TABLES[m.group(2)] = vals.reshape(dims)
    return _TABLES[name]


def pb_list(geom, n_refs=2, seed=2024, dst_pics=None, mix=None):
    """Prediction-block records for every picture of the destination ring.

    The picture is tiled by CUs of 4..64 samples per side (one shape per 64x64 region); each CU draws a
    mode: uni / bi (avg), DMVR+BDOF, BDOF only, BCW, explicit weights, affine (+PROF) or GPM, and a
    motion vector = smooth field (up to +-48 px) + jitter; 5 % of the CUs get vectors far outside the
    picture so the reference-window clamps are exercised.  CUs are cut into records of at most 16x16
    (4x4 + one chroma record per 8x8 for affine CUs), which is how the reference itself walks DMVR/BDOF
    and affine sub-blocks (vvc_inter.c:782-873).
    Returns (pbs, wp, prof).
    """
    rng = LCG(seed)
    mix = mix or dict(bi=70, dmvr=50, bdof_only=5, bcw=10, wp=5, affine=5, gpm=3)
    pics = range(geom.batch) if dst_pics is None else dst_pics
    n_wp = 8
    wp = np.zeros(n_wp, dtype=abi.WP_DTYPE)
    wp["weight"][:] = (rng.below(n_wp * 6, 256) - 128).reshape(n_wp, 2, 3)
    wp["offset"][:] = (rng.below(n_wp * 6, 256) - 128).reshape(n_wp, 2, 3)
    wp["log2_denom"][:] = rng.below(n_wp * 2, 8).reshape(n_wp, 2)
    ang_idx, ang_w, ang_mirror = spec_table("gpm_angle_idx"), spec_table("gpm_angle_to_weights_idx"), spec_table("gpm_angle_to_mirror")
    offx, offy = spec_table("gpm_weights_offset_x"), spec_table("gpm_weights_offset_y")
    recs, profs = [], []
    n_prof = 0
    W, H = geom.width, geom.height
    for k in pics:
        g64h, g64w = (H + 63) // 64, (W + 63) // 64
        shape_w = np.array([2, 3, 3, 4, 4, 4, 5, 5, 6, 6])[rng.below(g64h * g64w, 10)].reshape(g64h, g64w)
        shape_h = np.array([2, 3, 3, 4, 4, 4, 5, 5, 6, 6])[rng.below(g64h * g64w, 10)].reshape(g64h, g64w)
        # partial 64x64 regions at the right / bottom picture border: cap the CU size so the region's
        # visible part is tiled completely (every sample of the picture is predicted)
        def cap(rem):
            low = int(rem) & -int(rem)
            return 6 if rem >= 64 else max(2, low.bit_length() - 1)
        cap_w = np.array([cap(min(64, W - x * 64)) for x in range(g64w)])
        cap_h = np.array([cap(min(64, H - y * 64)) for y in range(g64h)])
        shape_w = np.minimum(shape_w, cap_w[None, :])
        shape_h = np.minimum(shape_h, cap_h[:, None])
        shape_h = np.where((shape_w == 2) & (shape_h == 2), 3, shape_h)        # no 4x4 inter CUs
        for lw in range(2, 7):
            for lh in range(2, 7):
                ry, rx = np.nonzero((shape_w == lw) & (shape_h == lh))
                if not len(ry):
                    continue
                cw, ch = 1 << lw, 1 << lh
                # every CU of these 64x64 regions
                iy, ix = np.mgrid[0:64 // ch, 0:64 // cw]
                cx = (rx[:, None] * 64 + ix.reshape(-1)[None, :] * cw).reshape(-1)
                cy = (ry[:, None] * 64 + iy.reshape(-1)[None, :] * ch).reshape(-1)
                ok = (cx + cw <= W) & (cy + ch <= H)
                cx, cy = cx[ok], cy[ok]
                n = len(cx)
                if not n:
                    continue
                r = rng.below(n, 100)
                small = cw == 4 or ch == 4
                bi = (r < mix["bi"]) & (not small)
                r2 = rng.below(n, 100)
                eligible = bi & (cw >= 8) & (ch >= 8) & (cw * ch >= 128)
                e0 = mix["dmvr"]; e1 = e0 + mix["bdof_only"]; e2 = e1 + mix["bcw"]; e3 = e2 + mix["wp"]
                dmvr = eligible & (r2 < e0)
                bdof = eligible & (r2 < e1)
                bcw = bi & (r2 >= e1) & (r2 < e2)
                wpf = (r2 >= e2) & (r2 < e3)                                   # uni or bi
                r3 = rng.below(n, 100)
                affine = (r3 < mix["affine"]) & ~dmvr & ~bdof & (cw >= 16) & (ch >= 16)
                gpm = (r3 >= mix["affine"]) & (r3 < mix["affine"] + mix["gpm"]) & (cw >= 8) & (ch >= 8) & ~affine
                dmvr &= ~gpm; bdof &= ~gpm; bcw &= ~gpm; wpf &= ~gpm
                pred = np.where(bi | gpm, abi.PF_BI, np.where(rng.below(n, 2) == 0, abi.PF_L0, abi.PF_L1))
                # motion: smooth field + jitter, 1/16 sample units
                ph = 0.37 * k
                fx = 768 * np.sin(0.0021 * cx + 0.0013 * cy + ph) + (rng.below(n, 129) - 64)
                fy = 512 * np.cos(0.0017 * cx - 0.0029 * cy + ph) + (rng.below(n, 129) - 64)
                mv = np.zeros((n, 2, 2), dtype=np.int64)
                mv[:, 0, 0], mv[:, 0, 1] = fx, fy
                mv[:, 1, 0] = -fx + (rng.below(n, 65) - 32)                     # roughly mirrored second list
                mv[:, 1, 1] = -fy + (rng.below(n, 65) - 32)
                ints = rng.below(n, 8) == 0                                      # some integer / half-pel vectors
                mv[ints] &= ~15
                far = rng.below(n, 20) == 0
                mv[far] += (rng.below(int(far.sum()) * 4, 2).reshape(-1, 2, 2) * 2 - 1) * (16 * max(W, H))
                refs = rng.below(n * 2, n_refs).reshape(n, 2)
                filt = np.where(affine, 2, (rng.below(n, 10) == 0).astype(np.int64))
                bcw_idx = np.where(bcw, rng.below(n, 4) + 1, 0)
                wp_idx = rng.below(n, n_wp)
                # GPM geometry (pred_gpm_blk, vvc_inter.c:466-505)
                part = rng.below(n, 64)
                ang = ang_idx[part]
                mirror = ang_mirror[ang]
                ox_ = offx[part, max(lh - 3, 0), max(lw - 3, 0)]
                oy_ = offy[part, max(lh - 3, 0), max(lw - 3, 0)]
                gsx = np.where(mirror == 1, -1, 1)
                gsy = np.where(mirror == 2, -112, 112)
                gbase = ang_w[ang] * 112 * 112 + np.where(mirror == 0, oy_ * 112 + ox_,
                                                          np.where(mirror == 1, oy_ * 112 + 111 - ox_, (111 - oy_) * 112 + ox_))
                # PROF tables for the affine CUs
                aff_ids = np.nonzero(affine)[0]
                prof_idx = np.zeros(n, dtype=np.int64)
                prof_idx[aff_ids] = n_prof + np.arange(len(aff_ids))
                if len(aff_ids):
                    pr = np.zeros(len(aff_ids), dtype=abi.PROF_DTYPE)
                    pr["diff_mv_x"][:] = (rng.below(len(aff_ids) * 32, 63) - 31).reshape(-1, 2, 16)
                    pr["diff_mv_y"][:] = (rng.below(len(aff_ids) * 32, 63) - 31).reshape(-1, 2, 16)
                    profs.append(pr)
                    n_prof += len(aff_ids)
                prof_flags = np.where(affine, rng.below(n, 4), 0)                # cb_prof_flag[0] | cb_prof_flag[1] << 1
                flags = (dmvr * abi.PB_DMVR) | (bdof * abi.PB_BDOF) | (gpm * abi.PB_GPM) | (wpf * abi.PB_WEIGHTED) \
                    | ((prof_flags & 1) * abi.PB_PROF0) | ((prof_flags >> 1) * abi.PB_PROF1)

                def expand(sel, tw, th, planes, mvs=None, sub=None):
                    ids = np.nonzero(sel)[0]
                    if not len(ids):
                        return
                    ty, tx = np.mgrid[0:ch // th, 0:cw // tw]
                    tx, ty = tx.reshape(-1) * tw, ty.reshape(-1) * th
                    m = len(tx)
                    rec = np.zeros(len(ids) * m, dtype=abi.PB_DTYPE)
                    rid = np.repeat(ids, m)
                    px, py = np.tile(tx, len(ids)), np.tile(ty, len(ids))
                    rec["x0"], rec["y0"] = cx[rid] + px, cy[rid] + py
                    rec["w"], rec["h"], rec["planes"], rec["pic"] = tw, th, planes, k
                    rec["pred_flag"], rec["ref"], rec["flags"] = pred[rid], refs[rid], flags[rid]
                    rec["filt"], rec["bcw_idx"], rec["wp"], rec["prof"] = filt[rid], bcw_idx[rid], wp_idx[rid], prof_idx[rid]
                    rec["gpm_step_x"], rec["gpm_step_y"] = gsx[rid], gsy[rid]
                    rec["gpm_weights"] = gbase[rid] + py * gsy[rid] + px * gsx[rid]
                    base_mv = mv[rid]
                    if sub is not None:     # affine: per sub-block vectors = CU vector + a linear field
                        gxx, gxy, gyx, gyy = sub
                        d = np.zeros_like(base_mv)
                        d[:, :, 0] = (gxx[rid] * px + gxy[rid] * py)[:, None] >> 4
                        d[:, :, 1] = (gyx[rid] * px + gyy[rid] * py)[:, None] >> 4
                        base_mv = base_mv + d
                    rec["mv"] = base_mv if mvs is None else mvs(base_mv, rid, px, py)
                    recs.append(rec)

                both = abi.PB_LUMA | abi.PB_CHROMA if geom.chroma_format_idc else abi.PB_LUMA
                expand(~affine, min(cw, 16), min(ch, 16), both)
                if affine.any():
                    sub = tuple(rng.below(n, 33) - 16 for _ in range(4))
                    expand(affine, 4, 4, abi.PB_LUMA, sub=sub)
                    if geom.chroma_format_idc:
                        def mvc(base, rid, px, py):
                            # derive_affine_mvc (vvc_inter.c:813-826): mv(x, y) + mv(x + 4, y + 4), rounded by 1
                            gxx, gxy, gyx, gyy = sub
                            d2 = np.zeros_like(base)
                            d2[:, :, 0] = (gxx[rid] * (px + 4) + gxy[rid] * (py + 4))[:, None] >> 4
                            d2[:, :, 1] = (gyx[rid] * (px + 4) + gyy[rid] * (py + 4))[:, None] >> 4
                            d1 = np.zeros_like(base)
                            d1[:, :, 0] = (gxx[rid] * px + gxy[rid] * py)[:, None] >> 4
                            d1[:, :, 1] = (gyx[rid] * px + gyy[rid] * py)[:, None] >> 4
                            s = (base + d1) + (base + d2)
                            return (s + 1 - (s >= 0)) >> 1
                        # chroma of affine CUs: one record per 8x8 luma, no PROF / weights from the luma flags kept
                        saved = flags.copy()
                        flags &= ~(abi.PB_PROF0 | abi.PB_PROF1)
                        ids_before = len(recs)
                        expand(affine, 8, 8, abi.PB_CHROMA, mvs=mvc)
                        flags[:] = saved
                        del ids_before
    pbs = np.concatenate(recs) if recs else np.zeros(0, dtype=abi.PB_DTYPE)
    prof = np.concatenate(profs) if profs else np.zeros(1, dtype=abi.PROF_DTYPE)
    # uni records carry no bi-only tools
    uni = pbs["pred_flag"] != abi.PF_BI
    pbs["bcw_idx"][uni] = 0
    return pbs, wp, prof


# ---------------------------------------------------------------------------------------------
# Intra leaf predictor records + their reference lines, and CIIP blocks.
# ---------------------------------------------------------------------------------------------
INTRA_ANGLES = np.array([0, 1, 2, 3, 4, 6, 8, 10, 12, 14, 16, 18, 20, 23, 26, 29, 32, 35, 39, 45, 51, 57, 64, 73, 86, 102,
                         128, 171, 256, 341, 512])


def intra_angle(mode):
    """ff_vvc_intra_pred_angle_derive (libavcodec/vvc/vvc_intra.c:661-681)."""
    idx = mode - 50 if mode > 34 else (18 - mode if mode > 0 else 16 - mode)
    return int(np.sign(idx)) * int(INTRA_ANGLES[abs(idx)]) if idx else 0


def intra_nscale(w, h, mode):
    """ff_vvc_nscale_derive for angular modes (vvc_intra.c:538-555); the float of inv_angle_derive is exact here."""
    angle = intra_angle(mode)
    inv = (32768 + angle) // (2 * angle)
    side = h if mode >= 50 else w
    return min(2, int(np.log2(side)) - int(np.floor(np.log2(3 * inv - 2))) + 8), inv


def intra_list(geom, seed=606):
    """Tile every plane with blocks and give each a random leaf predictor (planar, DC, V, H, angular with
    reference-line index / filter / PDPC variants incl. wide angles, MIP) plus exactly the reference samples it
    may read, as random samples of the picture's bit depth.  Returns (pbs, edges)."""
    rng = LCG(seed)
    maxv = (1 << geom.bit_depth) - 1
    recs, edge_parts, at = [], [], 16
    edge_parts.append(np.zeros(16, dtype=np.uint16))
    for k in range(geom.batch):
        lw, lh = tb_partition(geom, rng, stop_p=0.4)
        uh, uw = lw.shape
        uy, ux = np.mgrid[0:uh, 0:uw]
        origin = ((ux * 4) % (1 << lw) == 0) & ((uy * 4) % (1 << lh) == 0)
        x0, y0, l2w, l2h = ux[origin] * 4, uy[origin] * 4, lw[origin], lh[origin]
        inside = (x0 + (1 << l2w) <= geom.width) & (y0 + (1 << l2h) <= geom.height)
        x0, y0, l2w, l2h = x0[inside], y0[inside], l2w[inside], l2h[inside]
        for c in range(3 if geom.chroma_format_idc else 1):
            sh = 1 if c else 0
            n = len(x0)
            kind = rng.below(n, 7)
            r_mode, r_ref, r_flt, r_pd, r_tr = rng.below(n, 1 << 16), rng.below(n, 3), rng.below(n, 2), rng.below(n, 4), rng.below(n, 2)
            for i in range(n):
                w, h = (1 << int(l2w[i])) >> sh, (1 << int(l2h[i])) >> sh
                if w < 4:
                    continue        # 2-wide chroma intra blocks do not exist in VVC (pred_dc / pred_h store 4 samples at a time)
                r = np.zeros(1, dtype=abi.INTRA_PB_DTYPE)
                kd = int(kind[i])
                if kd == abi.INTRA_MIP and (c or w < 4 or h < 4):
                    kd = abi.INTRA_PLANAR
                r["x0"], r["y0"], r["w"], r["h"], r["c_idx"], r["pic"] = x0[i] >> sh, y0[i] >> sh, w, h, c, k
                before = after_top = after_left = 8
                if kd == abi.INTRA_MIP:
                    size_id = 0 if (w == 4 and h == 4) else (1 if (w == 4 or h == 4 or (w == 8 and h == 8)) else 2)
                    r["mode"] = int(r_mode[i]) % (16, 8, 6)[size_id]
                    r["flags"] = abi.INTRA_MIP_TRANSPOSED if r_tr[i] else 0
                elif kd in (abi.INTRA_ANGULAR_V, abi.INTRA_ANGULAR_H):
                    if kd == abi.INTRA_ANGULAR_V:
                        mode = 34 + int(r_mode[i]) % 47                       # 34..80
                        mode = 51 if mode == 50 else mode
                    else:
                        mode = -14 + int(r_mode[i]) % 48                      # -14..33
                        mode = {0: 2, 1: 3, 18: 19}.get(mode, mode)
                    ref_idx = 0 if c else int(r_ref[i])
                    angle = intra_angle(mode)
                    pdpc = 0
                    if w >= 4 and h >= 4 and not ref_idx and not (18 < mode < 50) and r_pd[i] != 0:
                        ns, inv = intra_nscale(w, h, mode)
                        if ns >= 0:
                            pdpc = 1
                            side_reach = ((256 + min(max(w, h), 3 << ns) * inv) >> 9) + 2
                            after_top, after_left = after_top + side_reach, after_left + side_reach
                    r["mode"], r["ref_idx"], r["filter_flag"], r["flags"] = mode, ref_idx, int(r_flt[i]), pdpc
                    reach = (((max(w, h) + 1 + ref_idx) * abs(angle)) >> 5) + ref_idx + 6
                    before += reach + ref_idx + 2
                    after_top, after_left = after_top + reach, after_left + reach
                r["kind"] = kd
                n_top, n_left = before + w + after_top + 1, before + h + after_left + 1
                e = (rng.take(n_top + n_left) & np.uint32(maxv)).astype(np.uint16)
                r["top"], r["left"] = at + before, at + n_top + before
                at += n_top + n_left
                edge_parts.append(e)
                recs.append(r)
    return np.concatenate(recs), np.concatenate(edge_parts)


def intra_blk_list(geom, seed=808, cell=256):
    """Independent VVCCudaIntraBlk records for the function-level tests of intra_pred / intra_cclm_pred: one block per
    cell x cell luma area, placed so that no block's reference samples (up to 2 * 64 + 3 to the right / below, 3 above /
    left) touch another block, so the whole list is one wavefront.  Every mode (0..66, wide angles through the mapping),
    reference line 0..2, ISP-shaped and BDPCM blocks, MIP, chroma blocks, the three CCLM modes; availability counts as
    the reference's functions can return them at that position (picture / CTB borders, partial lines, nothing)."""
    rng = LCG(seed)
    hs, vs, ctb = geom.hshift, geom.vshift, geom.ctb_size
    recs = []
    cells = [(k, cx, cy) for k in range(geom.batch) for cy in range(0, geom.height, cell) for cx in range(0, geom.width, cell)]
    n = len(cells)
    R = lambda m: rng.below(n, m)
    r_kind, r_ox, r_oy, r_lw, r_lh, r_mode, r_ref = R(16), R(16), R(16), R(5), R(5), R(1 << 16), R(6)
    r_al, r_at, r_l, r_t, r_ul, r_isp, r_bd, r_tr, r_col = R(8), R(8), R(1 << 16), R(1 << 16), R(4), R(8), R(10), R(2), R(2)
    for i, (k, cx, cy) in enumerate(cells):
        kd = int(r_kind[i])
        chroma = kd in (10, 11, 12, 13, 14, 15) and geom.chroma_format_idc
        cclm = kd in (13, 14, 15) and geom.chroma_format_idc
        mipk = kd in (8, 9)
        ox = 0 if r_ox[i] == 0 else 4 + 4 * int(r_ox[i])        # luma offsets inside the cell: 0 or 8..64
        oy = 0 if r_oy[i] == 0 else 4 + 4 * int(r_oy[i])
        wl, hl = 4 << int(r_lw[i]), 4 << int(r_lh[i])            # luma size 4..64
        if chroma:
            wl, hl = max(wl, 8 << (hs - 1) if hs else 4), max(hl, 8 << (vs - 1) if vs else 4)
        x0l, y0l = cx + ox, cy + oy
        if x0l + wl > geom.width or y0l + hl > geom.height:
            x0l, y0l = cx, cy
            if x0l + wl > geom.width or y0l + hl > geom.height:
                continue
        sx, sy = (hs, vs) if chroma else (0, 0)
        x, y, w, h = x0l >> sx, y0l >> sy, wl >> sx, hl >> sy
        pw = geom.width >> sx
        r = np.zeros(1, dtype=abi.INTRA_BLK_DTYPE)
        r["x0"], r["y0"], r["w"], r["h"], r["pic"], r["cb_w"], r["cb_h"] = x, y, w, h, k, wl, hl
        r["c_idx"] = (1 + int(r_mode[i]) % 2 if not cclm else 1) if chroma else 0
        # availability the reference's functions can report here
        max_y = min(geom.height, ((y0l >> geom.ctb_log2) + 1) << geom.ctb_log2) >> sy
        max_x = min(geom.width, ((x0l >> geom.ctb_log2) + 1) << geom.ctb_log2) >> sx
        left_max = max_y - y if x > 0 else 0
        if x0l % ctb == 0:
            L = 0 if (r_al[i] == 0 or x == 0) else left_max
        else:
            L = left_max if r_al[i] >= 3 else 0 if r_al[i] == 0 else 2 * (int(r_l[i]) % (left_max // 2 + 1))
        if y == 0:
            T = 0
        elif y0l % ctb == 0:
            T = 0 if r_at[i] == 0 else (pw - x) if r_at[i] >= 3 else 1 + int(r_t[i]) % (pw - x)
        else:
            T = (max_x - x) if r_at[i] >= 3 else 0 if r_at[i] == 0 else 2 * (int(r_t[i]) % ((max_x - x) // 2 + 1))
        r["avail_left"], r["avail_top"] = min(L, 255), min(T, 255)
        flags = abi.INTRA_F_UP_LEFT if (x > 0 and y > 0 and r_ul[i] != 0) else 0
        if cclm:
            r["kind"], r["pred_mode"] = abi.INTRA_KIND_CCLM, 81 + (kd - 13)
            flags |= (abi.INTRA_F_LUMA_AVAIL_T if T > 0 else 0) | (abi.INTRA_F_LUMA_AVAIL_L if L > 0 else 0)
            flags |= abi.INTRA_F_COLLOCATED if r_col[i] else 0
        elif mipk:
            size_id = 0 if (w == 4 and h == 4) else (1 if (w == 4 or h == 4 or (w == 8 and h == 8)) else 2)
            r["kind"], r["pred_mode"] = abi.INTRA_KIND_MIP, int(r_mode[i]) % (16, 8, 6)[size_id]
            flags |= abi.INTRA_F_MIP_TRANSP if r_tr[i] else 0
        else:
            m = int(r_mode[i]) % 80
            r["pred_mode"] = m if m < 67 else (0, 1, 18, 50, 2, 34, 66, 0, 1, 18, 50, 0, 1)[m - 67]
            if not chroma:
                r["ref_idx"] = (0, 0, 0, 0, 1, 2)[int(r_ref[i])] if y0l % ctb else 0
                if r_isp[i] == 0 and r["ref_idx"][0] == 0 and wl * hl >= 32:     # a sub-partition of a wl x hl coding block
                    if r_tr[i] and hl >= 8:
                        h = hl // (4 if hl >= 16 and wl * hl > 32 else 2)
                    elif wl >= 8:
                        w = wl // (4 if wl >= 16 and wl * hl > 32 else 2)
                    else:
                        h = hl // 2
                    r["w"], r["h"] = w, h
                    flags |= abi.INTRA_F_ISP
            flags |= abi.INTRA_F_BDPCM if r_bd[i] == 0 else 0
        r["flags"] = flags
        recs.append(r)
    return np.concatenate(recs)


def _morton(x, y):
    m = np.zeros_like(x, dtype=np.int64)
    for b in range(12):
        m |= ((x >> b) & 1).astype(np.int64) << (2 * b)
        m |= ((y >> b) & 1).astype(np.int64) << (2 * b + 1)
    return m


def intra_picture(geom, seed=909):
    """An all-intra picture ring for vvc_cuda_intra_recon_frame: the transform blocks of tb_list() double as coding units
    (luma block + its two chroma blocks, 4:2:0), decoded CTU by CTU in raster order and in z-order inside a CTU.  Every
    luma block gets a regular mode (0..66, MRL away from CTB tops), MIP or a BDPCM flag; the chroma pair either one CCLM
    record (three modes) or two regular records.  Availability counts and flags are what that decoding order makes
    available (a neighbour sample is available when its coding unit comes earlier, inside the limits of
    ff_vvc_get_left_available / _top_available); a block's wave is one more than the latest wave among the blocks whose
    samples it reads, chroma after its own luma.

    Returns dict(blks, blk_end, tbs, tb_end, coeffs: sorted by wave; dec_blks, dec_blk_end, dec_tbs, dec_tb_end: the same
    records in decoding order, one luma and one chroma step per coding unit - what a CPU decoder does)."""
    assert geom.chroma_format_idc == 1 and geom.hshift == 1 and geom.vshift == 1
    rng = LCG(seed)
    # square blocks: the z-order of their origins is then the decoding order of a quadtree, in which everything left of a
    # block over its whole height and above it over its whole width comes earlier (with rectangles the z-order of origins
    # is not the order of any partition tree: a block could then read samples that are only written later)
    # ... of at least 8x8 luma samples, so that every coding unit has chroma blocks (as in a stream, where the chroma of
    # smaller luma blocks is coded with their 8x8 parent): no chroma sample stays unreconstructed next to a block that reads it
    tbs, coeffs = tb_list(geom, seed=seed + 1, extras=False, square=True, min_log2=3)
    keep = ~((tbs["c_idx"] > 0) & ((tbs["log2_w"] < 2) | (tbs["log2_h"] < 2)))        # no 2-wide chroma blocks in VVC intra
    tbs = tbs[keep]
    ctb, cl2 = geom.ctb_size, geom.ctb_log2
    luma = np.nonzero(tbs["c_idx"] == 0)[0]
    lx, ly = tbs["x0"][luma].astype(np.int64), tbs["y0"][luma].astype(np.int64)
    key = (tbs["pic"][luma].astype(np.int64) << 40) | (((ly >> cl2) * geom.ctb_cols + (lx >> cl2)) << 24) | _morton(lx & (ctb - 1), ly & (ctb - 1))
    luma = luma[np.argsort(key, kind="stable")]
    # chroma TBs of a coding unit, by (pic, x, y)
    chroma_of = {}
    for i in np.nonzero(tbs["c_idx"] > 0)[0]:
        chroma_of.setdefault((int(tbs["pic"][i]), int(tbs["x0"][i]), int(tbs["y0"][i])), []).append(int(i))
    H, W = geom.height, geom.width
    order = [np.full((geom.batch, H >> s, W >> s), -1, dtype=np.int64) for s in (0, 1)]
    wave = [np.full((geom.batch, H >> s, W >> s), -1, dtype=np.int64) for s in (0, 1)]
    n = len(luma)
    R = lambda m: rng.below(n, m)
    r_kind, r_mode, r_ref, r_bd, r_tr, r_ck, r_cm, r_col = R(10), R(1 << 16), R(6), R(12), R(2), R(5), R(1 << 16), R(2)
    blks, blk_wave, tb_wave = [], [], np.zeros(len(tbs), dtype=np.int64)
    dec_blks, dec_blk_end, dec_tb_idx, dec_tb_end, dec_wave, dec_ctu = [], [], [], [], [], []

    def avail(ch, k, o, x, y, x0l, y0l):
        """(left count, top count, up-left) of the block at plane position (x, y) of plane type ch"""
        sh = ch
        om = order[ch][k]
        ph, pw = om.shape
        max_y = min(H, ((y0l >> cl2) + 1) << cl2) >> sh
        max_x = min(W, ((x0l >> cl2) + 1) << cl2) >> sh
        L = T = 0
        if x > 0:
            col = om[y:max_y, x - 1]
            ok = (col >= 0) & (col < o)
            L = int(len(ok) if ok.all() else np.argmin(ok))
        if y > 0:
            row = om[y - 1, x:(pw if y0l % ctb == 0 else max_x)]
            ok = (row >= 0) & (row < o)
            T = int(len(ok) if ok.all() else np.argmin(ok))
        ul = x > 0 and y > 0 and 0 <= om[y - 1, x - 1] < o
        return min(L, 255), min(T, 255), ul

    def deps(ch, k, x, y, w, h):
        """latest wave among the samples the block at (x, y, w, h) of plane type ch may read"""
        wm = wave[ch][k]
        ya, xa = max(y - 3, 0), max(x - 3, 0)
        m = -1
        if y > 0:
            m = max(m, int(wm[ya:y, xa:x + 2 * w + 4].max()))
        if x > 0:
            m = max(m, int(wm[ya:y + 2 * h + 4, xa:x].max()))
        return m

    for j, ti in enumerate(luma):
        t = tbs[ti]
        k, x0, y0, w, h = int(t["pic"]), int(t["x0"]), int(t["y0"]), 1 << int(t["log2_w"]), 1 << int(t["log2_h"])
        o = j
        # ---- luma block ----
        r = np.zeros(1, dtype=abi.INTRA_BLK_DTYPE)
        r["x0"], r["y0"], r["w"], r["h"], r["pic"], r["cb_w"], r["cb_h"] = x0, y0, w, h, k, w, h
        L, T, ul = avail(0, k, o, x0, y0, x0, y0)
        r["avail_left"], r["avail_top"] = L, T
        flags = abi.INTRA_F_UP_LEFT if ul else 0
        if r_kind[j] < 2:
            size_id = 0 if (w == 4 and h == 4) else (1 if (w == 4 or h == 4 or (w == 8 and h == 8)) else 2)
            r["kind"], r["pred_mode"] = abi.INTRA_KIND_MIP, int(r_mode[j]) % (16, 8, 6)[size_id]
            flags |= abi.INTRA_F_MIP_TRANSP if r_tr[j] else 0
        else:
            m = int(r_mode[j]) % 75
            r["pred_mode"] = m if m < 67 else (0, 1, 18, 50, 0, 1, 34, 66)[m - 67]
            r["ref_idx"] = (0, 0, 0, 0, 1, 2)[int(r_ref[j])] if y0 % ctb else 0
            flags |= abi.INTRA_F_BDPCM if r_bd[j] == 0 else 0
        r["flags"] = flags
        wl = deps(0, k, x0, y0, w, h) + 1
        order[0][k, y0:y0 + h, x0:x0 + w] = o
        wave[0][k, y0:y0 + h, x0:x0 + w] = wl
        blks.append(r); blk_wave.append(wl)
        tb_wave[ti] = wl
        dec_blks.append(r); dec_blk_end.append(len(dec_blks)); dec_tb_idx.append(ti); dec_tb_end.append(len(dec_tb_idx))
        dec_wave.append(wl); dec_ctu.append((k, x0 // ctb, y0 // ctb))
        # ---- its chroma blocks ----
        ctbs = chroma_of.get((k, x0 >> 1, y0 >> 1), [])
        if w < 8 or h < 8:
            assert not ctbs
            continue
        x, y, cw, chh = x0 >> 1, y0 >> 1, w >> 1, h >> 1
        L, T, ul = avail(1, k, o, x, y, x0, y0)
        wc = max(deps(1, k, x, y, cw, chh), wl) + 1
        recs = []
        if r_ck[j] < 2:
            q = np.zeros(1, dtype=abi.INTRA_BLK_DTYPE)
            q["x0"], q["y0"], q["w"], q["h"], q["pic"], q["c_idx"], q["cb_w"], q["cb_h"] = x, y, cw, chh, k, 1, w, h
            q["kind"], q["pred_mode"] = abi.INTRA_KIND_CCLM, 81 + int(r_cm[j]) % 3
            q["avail_left"], q["avail_top"] = L, T
            lo = order[0][k]
            at = y0 > 0 and 0 <= lo[y0 - 1, x0] < o
            al = x0 > 0 and 0 <= lo[y0, x0 - 1] < o
            q["flags"] = (abi.INTRA_F_UP_LEFT if ul else 0) | (abi.INTRA_F_LUMA_AVAIL_T if at else 0) | (abi.INTRA_F_LUMA_AVAIL_L if al else 0) | \
                         (abi.INTRA_F_COLLOCATED if r_col[j] else 0)
            # the luma it reads: above / left of the coding unit as far as the T / L modes reach, and the unit itself
            lw_ = wave[0][k]
            ya, xa = max(y0 - 3, 0), max(x0 - 3, 0)
            m = wl
            if y0 > 0:
                m = max(m, int(lw_[ya:y0, xa:x0 + 2 * w + 4].max()))
            if x0 > 0:
                m = max(m, int(lw_[ya:y0 + 2 * h + 4, xa:x0].max()))
            wc = max(wc, m + 1)
            recs.append(q)
        else:
            m = int(r_cm[j]) % 75
            pm = m if m < 67 else (0, 1, 18, 50, 0, 1, 34, 66)[m - 67]
            for c in (1, 2):
                q = np.zeros(1, dtype=abi.INTRA_BLK_DTYPE)
                q["x0"], q["y0"], q["w"], q["h"], q["pic"], q["c_idx"], q["cb_w"], q["cb_h"] = x, y, cw, chh, k, c, w, h
                q["pred_mode"], q["avail_left"], q["avail_top"], q["flags"] = pm, L, T, abi.INTRA_F_UP_LEFT if ul else 0
                recs.append(q)
        order[1][k, y:y + chh, x:x + cw] = o
        wave[1][k, y:y + chh, x:x + cw] = wc
        for q in recs:
            blks.append(q); blk_wave.append(wc); dec_blks.append(q)
        dec_blk_end.append(len(dec_blks))
        for ci in ctbs:
            tb_wave[ci] = wc
            dec_tb_idx.append(ci)
        dec_tb_end.append(len(dec_tb_idx))
        dec_wave.append(wc); dec_ctu.append((k, x0 // ctb, y0 // ctb))
    blks = np.concatenate(blks)
    blk_wave = np.array(blk_wave)
    n_waves = int(max(blk_wave.max(), tb_wave.max())) + 1
    bo, to = np.argsort(blk_wave, kind="stable"), np.argsort(tb_wave, kind="stable")
    return dict(blks=blks[bo], blk_end=np.cumsum(np.bincount(blk_wave, minlength=n_waves)).astype(np.int32),
                tbs=tbs[to], tb_end=np.cumsum(np.bincount(tb_wave, minlength=n_waves)).astype(np.int32), coeffs=coeffs,
                dec_blks=np.concatenate(dec_blks), dec_blk_end=np.array(dec_blk_end, dtype=np.int32),
                dec_tbs=tbs[np.array(dec_tb_idx)], dec_tb_end=np.array(dec_tb_end, dtype=np.int32), n_waves=n_waves,
                dec_wave=np.array(dec_wave, dtype=np.int64), dec_ctu=np.array(dec_ctu, dtype=np.int64))


def intra_step_order(case, how):
    """The steps of intra_picture() in another legal order for vvc_cuda_intra_recon_frame_ordered (every step after the steps
    whose samples it reads): "decode" = as decoded, picture after picture; "ctu_wavefront" = by CTU anti-diagonal x + 2 y
    (the order of wavefront parallel processing, which a decoder knows without analysing blocks), pictures interleaved;
    "block_wave" = by the block's own wave.  Returns (blks, blk_end, tbs, tb_end)."""
    n = len(case["dec_blk_end"])
    if how == "decode":
        perm = np.arange(n)
    elif how == "ctu_wavefront":
        k, cx, cy = case["dec_ctu"].T
        perm = np.lexsort((np.arange(n), k, cx + 2 * cy))
    elif how == "block_wave":
        perm = np.argsort(case["dec_wave"], kind="stable")
    else:
        raise ValueError(how)

    def gather(items, ends):
        starts = np.concatenate(([0], ends[:-1])).astype(np.int64)
        lens = (ends - starts)[perm]
        idx = np.repeat(starts[perm] - np.concatenate(([0], np.cumsum(lens)[:-1])), lens) + np.arange(int(lens.sum()))
        return items[idx], np.cumsum(lens).astype(np.int32)
    blks, blk_end = gather(case["dec_blks"], case["dec_blk_end"])
    tbs, tb_end = gather(case["dec_tbs"], case["dec_tb_end"])
    return blks, blk_end, tbs, tb_end


def ciip_list(geom, seed=707):
    """CIIP blocks on a random partition: every block of every plane with intra weight 1..3."""
    rng = LCG(seed)
    recs = []
    for k in range(geom.batch):
        lw, lh = tb_partition(geom, rng, stop_p=0.4)
        uh, uw = lw.shape
        uy, ux = np.mgrid[0:uh, 0:uw]
        origin = ((ux * 4) % (1 << lw) == 0) & ((uy * 4) % (1 << lh) == 0)
        x0, y0, l2w, l2h = ux[origin] * 4, uy[origin] * 4, lw[origin], lh[origin]
        inside = (x0 + (1 << l2w) <= geom.width) & (y0 + (1 << l2h) <= geom.height)
        x0, y0, l2w, l2h = x0[inside], y0[inside], l2w[inside], l2h[inside]
        wi = rng.below(len(x0), 3) + 1
        for c in range(3 if geom.chroma_format_idc else 1):
            sh = 1 if c else 0
            r = np.zeros(len(x0), dtype=abi.CIIP_DTYPE)
            r["x0"], r["y0"], r["w"], r["h"] = x0 >> sh, y0 >> sh, (1 << l2w) >> sh, (1 << l2h) >> sh
            r["c_idx"], r["pic"], r["intra_weight"] = c, k, wi
            recs.append(r)
    return np.concatenate(recs)
