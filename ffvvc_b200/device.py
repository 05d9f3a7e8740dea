"""Device-memory plumbing on top of PyTorch (allocation, streams, events only)."""
import numpy as np
import torch

from . import abi


class DeviceFrames:
    """A picture ring resident in HBM: one uint16 tensor per plane, 256-byte row pitch."""

    def __init__(self, geom, device="cuda:0", planes=None):
        self.geom = geom
        self.t = []
        for c in range(3):
            shape = geom.plane_shape(c)
            if planes is not None:
                t = torch.from_numpy(planes[c].view(np.int16)).to(device)
            else:
                t = torch.zeros(shape, dtype=torch.int16, device=device)
            self.t.append(t)
        self.desc = abi.frame_desc(
            geom, [t.data_ptr() for t in self.t],
            [t.stride(1) * 2 for t in self.t], [t.stride(0) * 2 for t in self.t])

    def to_numpy(self):
        return [t.cpu().numpy().view(np.uint16) for t in self.t]

    @property
    def nbytes(self):
        return sum(t.numel() * 2 for t in self.t)


def to_device(arr, device="cuda:0"):
    """Upload a numpy array of any dtype as raw bytes; returns (tensor, device pointer)."""
    raw = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
    t = torch.from_numpy(raw.copy()).to(device)
    return t, t.data_ptr()
