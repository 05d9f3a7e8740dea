"""vvc_cuda_notify: the stream-ordered completion report a decoder uses in place of blocking in vvc_cuda_sync
(report_frame_progress, libavcodec/vvc/vvc_thread.c:390-410)."""
import threading
import time

import numpy as np
import pytest

from ffvvc_b200 import abi, synth


@pytest.mark.gpu
def test_notify_fires_in_submission_order_after_the_work():
    import torch
    from ffvvc_b200 import device, lib
    geom = abi.FrameGeom(3840, 2160, batch=4)
    _, inv = synth.lmcs_luts(10)
    enable = np.ones(geom.ctb_count * geom.batch, np.uint8)
    ctx = lib.Context(0)
    fired, done = [], threading.Event()
    try:
        with torch.cuda.stream(ctx.torch_stream()):
            fr = device.DeviceFrames(geom, planes=synth.uniform_planes(geom, seed=5))
            keep = [device.to_device(inv), device.to_device(enable)]
            ctx.sync()
            for _ in range(40):
                ctx.lmcs_frame(fr.desc, keep[0][1], keep[1][1])
            ctx.notify(lambda status: fired.append((1, status, time.perf_counter())))
            for _ in range(40):
                ctx.lmcs_frame(fr.desc, keep[0][1], keep[1][1])
            ctx.notify(lambda status: (fired.append((2, status, time.perf_counter())), done.set()))
            assert done.wait(30.0), "the second report never arrived"
            ctx.sync()
        assert [f[0] for f in fired] == [1, 2] and all(f[1] == 0 for f in fired)
        assert fired[0][2] <= fired[1][2]
    finally:
        ctx.close()


@pytest.mark.gpu
def test_notify_rejects_null_and_reports_sticky_errors():
    from ffvvc_b200 import lib
    ctx = lib.Context(0)
    try:
        assert ctx.lib.vvc_cuda_notify(ctx.handle, lib.NOTIFY_FN(0), None) == -2
        got = []
        ctx.notify(lambda status: got.append(status))
        ctx.sync()
        assert got == [0]
    finally:
        ctx.close()
