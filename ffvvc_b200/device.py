"""Device-memory plumbing on top of PyTorch (allocation, streams, events only)."""
import numpy as np
import torch

from . import abi


class DeviceFrames:
    """A picture ring resident in HBM: one uint16 tensor per plane, 256-byte row pitch."""

    def __init__(self, geom, device="cuda:0", planes=None, pad=0):
        """pad: luma samples of margin allocated around every plane (the pre-padded DPB format of
        VVC_CUDA_OPT_REF_PAD; the descriptor still points at sample (0, 0))."""
        self.geom, self.pad = geom, pad
        self.t, self.margin = [], []
        for c in range(3):
            b, h, pitch = geom.plane_shape(c)
            mx, my = (pad >> geom.hshift, pad >> geom.vshift) if c else (pad, pad)
            if pad:
                pitch = (geom.plane_wh(c)[0] + 2 * mx + 127) // 128 * 128
                t = torch.zeros((b, h + 2 * my, pitch), dtype=torch.int16, device=device)
                if planes is not None:
                    src = torch.from_numpy(planes[c].view(np.int16)).to(device)
                    t[:, my:my + h, mx:mx + src.shape[2]] = src[:, :, :min(src.shape[2], pitch - mx)]
            elif planes is not None:
                t = torch.from_numpy(planes[c].view(np.int16)).to(device)
            else:
                t = torch.zeros((b, h, pitch), dtype=torch.int16, device=device)
            self.t.append(t)
            self.margin.append((mx, my))
        self.desc = abi.frame_desc(
            geom, [t.data_ptr() + (my * t.stride(1) + mx) * 2 for t, (mx, my) in zip(self.t, self.margin)],
            [t.stride(1) * 2 for t in self.t], [t.stride(0) * 2 for t in self.t])

    def to_numpy(self, with_margin=False):
        full = [t.cpu().numpy().view(np.uint16) for t in self.t]
        if with_margin or not self.pad:
            return full
        return [np.ascontiguousarray(a[:, my:a.shape[1] - my, mx:]) for a, (mx, my) in zip(full, self.margin)]

    @property
    def nbytes(self):
        return sum(t.numel() * 2 for t in self.t)


def to_device(arr, device="cuda:0"):
    """Upload a numpy array of any dtype as raw bytes; returns (tensor, device pointer)."""
    raw = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
    t = torch.from_numpy(raw.copy()).to(device)
    return t, t.data_ptr()
