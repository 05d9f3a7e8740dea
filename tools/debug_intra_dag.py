import sys, os, ctypes as C, numpy as np
sys.path.insert(0, os.getcwd())
import torch
from ffvvc_b200 import abi, synth, lib, device
from tests import util
w,h,batch,seed = 832,480,1,5
geom = abi.FrameGeom(w, h, batch=batch)
case = synth.intra_picture(geom, seed=seed)
planes = abi.alloc_planes(geom, fill=512)
want = [a.copy() for a in planes]
co = case["coeffs"].copy()
cd = abi.coeffs_desc(co.ctypes.data, co.size)
util.oracle().vvco_intra_recon_frame(abi.frame_from_numpy(geom, want), case["dec_blks"].ctypes.data, case["dec_blk_end"].ctypes.data, C.byref(cd), case["dec_tbs"].ctypes.data, case["dec_tb_end"].ctypes.data, len(case["dec_blk_end"]), 15)
ctx = lib.Context(0)
torch.cuda.set_stream(ctx.torch_stream())
keep = [device.to_device(a) for a in (case["dec_blks"], case["dec_blk_end"], case["coeffs"], case["dec_tbs"], case["dec_tb_end"])]
for g in sys.argv[1:]:
    os.environ["VVC_CUDA_INTRA_GRID"] = g
    fr = device.DeviceFrames(geom, planes=planes)
    ctx.intra_recon_frame_ordered(fr.desc, keep[0][1], keep[1][1], abi.coeffs_desc(keep[2][1], case["coeffs"].size), keep[3][1], keep[4][1], len(case["dec_blk_end"]), len(case["dec_blks"]), len(case["dec_tbs"]), 15)
    ctx.sync()
    got = fr.to_numpy()
    out = []
    for c in range(3):
        wv = geom.plane_wh(c)[0]
        bad = np.argwhere(got[c][:, :, :wv] != want[c][:, :, :wv])
        out.append((len(bad), tuple(bad[0]) if len(bad) else None))
    print("grid", g, out, flush=True)
    if out[1][0]:
        k,y,x = out[1][1]
        blks = case["dec_blks"]
        hit = np.nonzero((blks["c_idx"]>0)&(blks["x0"]<=x)&(x<blks["x0"]+blks["w"].astype(int))&(blks["y0"]<=y)&(y<blks["y0"]+blks["h"].astype(int)))[0]
        print("   first bad chroma block(s):", [(int(i), blks[i]) for i in hit])
