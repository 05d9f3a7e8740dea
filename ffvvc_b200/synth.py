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


def alf_params(geom, seed=777, all_on=True):
    """Per-CTB ALF parameters + filter sets, SURVEY.md 8(d) config 1 distributions."""
    rng = LCG(seed)
    n = geom.ctb_count * geom.batch
    ctbs = np.zeros(n, dtype=abi.ALF_CTB_DTYPE)
    if all_on:
        ctbs["ctb_flag"][:] = 1
    else:
        ctbs["ctb_flag"][:] = rng.below(n * 3, 2).reshape(n, 3)
    ctbs["filt_set_idx_y"] = rng.below(n, 17)
    ctbs["chroma_alt_idx"][:] = rng.below(n * 2, 8).reshape(n, 2)
    ctbs["cc_idc"][:] = rng.below(n * 2, 5).reshape(n, 2)
    sets = np.zeros(1, dtype=abi.ALF_SETS_DTYPE)
    # APS luma: random int8-range coefficients (tests/checkasm/vvc_alf.c:55-59), clip idx rnd%4
    sets["luma_coeff"][0] = (rng.below(8 * 25 * 12, 256) - 128).reshape(8, 25, 12)
    sets["luma_clip_idx"][0] = rng.below(8 * 25 * 12, 4).reshape(8, 25, 12)
    sets["chroma_coeff"][0] = (rng.below(8 * 6, 256) - 128).reshape(8, 6)
    sets["chroma_clip_idx"][0] = rng.below(8 * 6, 4).reshape(8, 6)
    # CC-ALF coefficients are 0 or +-2^k, k <= 6 (vvc_ps.c:810-819)
    mag = rng.below(2 * 5 * 7, 8)
    sgn = rng.below(2 * 5 * 7, 2) * 2 - 1
    cc = np.where(mag == 0, 0, sgn * (1 << np.maximum(mag - 1, 0)))
    sets["cc_coeff"][0] = cc.reshape(2, 5, 7)
    return ctbs, sets
