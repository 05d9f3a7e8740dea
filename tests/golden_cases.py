"""Golden cases: a whole small picture pushed through every stage of the hot path in the reference's
stage order (INTER -> RECON residual -> LMCS -> DEBLOCK_V -> DEBLOCK_H -> SAO -> ALF,
libavcodec/vvc/vvc_thread.c:41-51), on seeded synthetic inputs.  tools/gen_golden.py runs the COMPILED
REFERENCE (oracle/_ref) over them and commits digests + small raw crops to tests/golden/; the tests hold
the oracle (CPU) and the CUDA library (GPU) to those files, so parity stays pinned on machines that
have no /root/reference."""
import ctypes as C
import hashlib

import numpy as np

from ffvvc_b200 import abi, synth

CASES = {
    # name: (width, height, bit_depth, ctb_log2, seed, uniform pixels)
    "cif_struct": (352, 288, 10, 7, 101, False),
    "wqvga_uniform": (416, 240, 10, 7, 202, True),
    "ctb64_struct": (256, 192, 10, 6, 303, False),
    "bd12_struct": (256, 128, 12, 7, 404, False),
}
STRESS_MIX = dict(bi=70, dmvr=35, bdof_only=15, bcw=15, wp=15, affine=20, gpm=15)


def build_case(name, lfnst_set_of):
    w, h, bd, ctb_log2, seed, uniform = CASES[name]
    geom = abi.FrameGeom(w, h, bit_depth=bd, ctb_log2=ctb_log2)
    gref = abi.FrameGeom(w, h, bit_depth=bd, ctb_log2=ctb_log2, batch=3)
    gen = synth.uniform_planes if uniform else synth.struct_planes
    case = dict(geom=geom, gref=gref, refs=gen(gref, seed=seed))
    case["pbs"], case["wp"], case["prof"] = synth.pb_list(geom, n_refs=3, seed=seed + 1, mix=STRESS_MIX)
    case["tbs"], case["coeffs"] = synth.tb_list(geom, seed=seed + 2, lfnst_set_of=lfnst_set_of, extras=False)
    # the same tiling as quantised levels in the 16-bit window layout, dequantised by the stage itself
    case["qtbs"], case["qwin"] = abi.pack_window16(synth.tb_for_window(case["tbs"]), case["coeffs"])
    case["quant"], case["sl"] = synth.tb_quant(case["qtbs"], seed=seed + 9, scaling=True)
    case["fwd"], case["inv"] = synth.lmcs_luts(bd, seed=seed + 3)
    case["maps"] = synth.deblock_maps(geom, seed=seed + 4, qp_base=27, qp_span=16)
    case["sao"] = synth.sao_params(geom, seed=seed + 5)
    case["alf"], case["sets"] = synth.alf_params(geom, seed=seed + 6)
    case["intra_pbs"], case["intra_edges"] = synth.intra_list(geom, seed=seed + 7)
    case["ciip"] = synth.ciip_list(geom, seed=seed + 8)
    return case


class HostBackend:
    """Stage functions of a CPU library (prefix 'vvco_' = oracle, 'vvcref_' = compiled reference)."""

    def __init__(self, lib, prefix):
        self.lib, self.prefix = lib, prefix

    def fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def inter(self, geom, gref, dst, refs, pbs, wp, prof, out):
        self.fn("inter_frame")(abi.frame_from_numpy(geom, dst), abi.frame_from_numpy(gref, refs), pbs.ctypes.data, len(pbs),
                               wp.ctypes.data, prof.ctypes.data, out.ctypes.data)

    def itx(self, geom, pic, coeffs, tbs, rng):
        self.fn("itx_frame")(abi.frame_from_numpy(geom, pic), coeffs.ctypes.data, tbs.ctypes.data, len(tbs), rng)

    def itx_q(self, geom, pic, win, tbs, quant, sl, rng):
        co = abi.coeffs_desc(win.ctypes.data, win.size, abi.COEFF_WINDOW16, quant.ctypes.data, sl.ctypes.data)
        fn = self.fn("itx_frame_q")
        fn.argtypes = [C.POINTER(abi.VVCCudaFrame), C.POINTER(abi.VVCCudaCoeffs), C.c_void_p, C.c_int, C.c_int]
        fn.restype = None
        fn(abi.frame_from_numpy(geom, pic), C.byref(co), tbs.ctypes.data, len(tbs), rng)

    def lmcs(self, geom, pic, lut):
        self.fn("lmcs_frame")(abi.frame_from_numpy(geom, pic), lut.ctypes.data, None)

    def deblock(self, geom, dst, src, maps):
        md = abi.deblock_maps_desc(geom, maps)
        tmp = abi.alloc_planes(geom)
        self.fn("deblock_frame")(abi.frame_from_numpy(geom, tmp), abi.frame_from_numpy(geom, src), C.byref(md), 1)
        self.fn("deblock_frame")(abi.frame_from_numpy(geom, dst), abi.frame_from_numpy(geom, tmp), C.byref(md), 0)

    def sao(self, geom, dst, src, ctbs):
        self.fn("sao_frame")(abi.frame_from_numpy(geom, dst), abi.frame_from_numpy(geom, src), ctbs.ctypes.data)

    def alf(self, geom, dst, src, ctbs, sets):
        self.fn("alf_frame")(abi.frame_from_numpy(geom, dst), abi.frame_from_numpy(geom, src), ctbs.ctypes.data, sets.ctypes.data, 0)

    def intra(self, geom, pic, pbs, edges):
        self.fn("intra_leaf_frame")(abi.frame_from_numpy(geom, pic), pbs.ctypes.data, len(pbs), edges.ctypes.data)

    def ciip(self, geom, dst, inter, blocks):
        self.fn("ciip_frame")(abi.frame_from_numpy(geom, dst), abi.frame_from_numpy(geom, inter), blocks.ctypes.data, len(blocks))


class CudaHostBackend:
    """The product library through its *_host C-ABI entries (host pointers in, host pointers out)."""

    def __init__(self, ctx):
        self.ctx = ctx

    def inter(self, geom, gref, dst, refs, pbs, wp, prof, out):
        self.ctx.inter_frame_host(abi.frame_from_numpy(geom, dst), abi.frame_from_numpy(gref, refs), pbs.ctypes.data, len(pbs),
                                  wp.ctypes.data, len(wp), prof.ctypes.data, len(prof), out.ctypes.data)

    def itx(self, geom, pic, coeffs, tbs, rng):
        self.ctx.itx_frame_host(abi.frame_from_numpy(geom, pic), coeffs.ctypes.data, len(coeffs), tbs.ctypes.data, len(tbs), rng)

    def itx_q(self, geom, pic, win, tbs, quant, sl, rng):
        co = abi.coeffs_desc(win.ctypes.data, win.size, abi.COEFF_WINDOW16, quant.ctypes.data, sl.ctypes.data)
        self.ctx.itx_frame_q_host(abi.frame_from_numpy(geom, pic), co, tbs.ctypes.data, len(tbs), rng)

    def lmcs(self, geom, pic, lut):
        self.ctx.lmcs_frame_host(abi.frame_from_numpy(geom, pic), lut.ctypes.data, None)

    def deblock(self, geom, dst, src, maps):
        md = abi.deblock_maps_desc(geom, maps)
        self.ctx.deblock_frame_host(abi.frame_from_numpy(geom, dst), abi.frame_from_numpy(geom, src), md)

    def sao(self, geom, dst, src, ctbs):
        self.ctx.sao_frame_host(abi.frame_from_numpy(geom, dst), abi.frame_from_numpy(geom, src), ctbs.ctypes.data)

    def alf(self, geom, dst, src, ctbs, sets):
        self.ctx.alf_frame_host(abi.frame_from_numpy(geom, dst), abi.frame_from_numpy(geom, src), ctbs.ctypes.data, sets.ctypes.data)

    def intra(self, geom, pic, pbs, edges):
        self.ctx.intra_leaf_frame_host(abi.frame_from_numpy(geom, pic), pbs.ctypes.data, len(pbs), edges.ctypes.data, len(edges))

    def ciip(self, geom, dst, inter, blocks):
        self.ctx.ciip_frame_host(abi.frame_from_numpy(geom, dst), abi.frame_from_numpy(geom, inter), blocks.ctypes.data, len(blocks))


def run_case(case, be):
    """Returns {stage: [arrays]} after each stage of the reconstruction."""
    g, out = case["geom"], {}
    vis = lambda planes: [p[:, :, :g.plane_wh(c)[0]].copy() for c, p in enumerate(planes)]
    pic = abi.alloc_planes(g, fill=1 << (g.bit_depth - 1))
    dm = np.zeros(len(case["pbs"]), dtype=abi.DMVR_OUT_DTYPE)
    be.inter(g, case["gref"], pic, case["refs"], case["pbs"], case["wp"], case["prof"], dm)
    is_dm = (case["pbs"]["flags"] & abi.PB_DMVR) != 0
    out["inter"] = vis(pic) + [np.ascontiguousarray(dm[is_dm])]
    # intra leaf predictors over a tiling of their own, then the CIIP blend of that picture with the inter one
    ipic = abi.alloc_planes(g, fill=3)
    be.intra(g, ipic, case["intra_pbs"], case["intra_edges"])
    out["intra"] = vis(ipic)
    be.ciip(g, ipic, [p.copy() for p in pic], case["ciip"])
    out["ciip"] = vis(ipic)
    qpic = [p.copy() for p in pic]
    be.itx_q(g, qpic, case["qwin"], case["qtbs"], case["quant"], case["sl"], 15)
    out["residual_q"] = vis(qpic)
    coeffs = case["coeffs"].copy()
    be.itx(g, pic, coeffs, case["tbs"], 15)
    out["residual"] = vis(pic)
    be.lmcs(g, pic, case["inv"])
    out["lmcs"] = vis(pic)
    a, b = abi.alloc_planes(g), abi.alloc_planes(g)
    be.deblock(g, a, pic, case["maps"])
    out["deblock"] = vis(a)
    be.sao(g, b, a, case["sao"])
    out["sao"] = vis(b)
    be.alf(g, a, b, case["alf"], case["sets"])
    out["alf"] = vis(a)
    return out


def digest(arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def crop(arrays):
    """A small raw crop (first 8 rows x 32 samples of every plane) stored verbatim next to the digest."""
    return [a[0, :8, :32].astype(np.int64).tolist() for a in arrays[:3]]
