"""Device-side clip driver: the per-frame loop of evaluate_originalSize() (reference
main_flownetS_pyramid_noprevloss_dataloader.py:535-630) behind one call per frame.

    net = get_net(); load_and_assign_npz_dict(...)
    stab = ClipStabilizer(net, n_clips=1, height=out_h, width=out_w)
    for i in range(total_frames):                       # main_dl.py:540
        ret, frame = cap.read()                         # :547  (BGR uint8)
        out.write(stab.step(frame))                     # :550-630 on the GPU

The history of stabilised frames lives on the device as a ring of resized uint8 slices; a step moves 3 bytes per
pixel each way over PCIe.  Several clips of the same size advance in lockstep as one batch (n_clips > 1).
"""
from __future__ import annotations

import collections
import ctypes as C

import numpy as np

from . import _lib


class ClipStabilizer:
    def __init__(self, net, n_clips=1, height=720, width=1280):
        self._lib = _lib.load()
        self.net = net                      # keeps the FlowNetSPyramid (and its ofs_net handle) alive
        self.n, self.h, self.w = int(n_clips), int(height), int(width)
        self._h = C.c_void_p()
        _lib.check(self._lib.ofs_clips_create(C.byref(self._h), net._h, self.n, self.h, self.w))
        net.retain()                        # the C side keeps the raw ofs_net*: the net must not be destroyed under it
        self._pending = collections.deque()   # (frames, out, outf, single) of submitted steps: keeps the buffers alive

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.ofs_clips_destroy(self._h)
            self._h = C.c_void_p()
            self.net.release()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def frame_index(self):
        return int(self._lib.ofs_clips_frame_index(self._h))

    def reset(self):
        _lib.check(self._lib.ofs_clips_reset(self._h))
        self._pending.clear()

    @property
    def depth(self):
        """How many submitted steps may be in flight."""
        return int(self._lib.ofs_clips_depth())

    @property
    def in_flight(self):
        return len(self._pending)

    def pinned_buffer(self):
        """A page-locked uint8 [n_clips,H,W,3] numpy array: frames decoded straight into it (and outputs written into
        one) cross PCIe at full speed; pageable arrays are staged by the driver at a fraction of it."""
        import torch

        return torch.empty((self.n, self.h, self.w, 3), dtype=torch.uint8).pin_memory().numpy()

    def _check_args(self, frames_bgr, return_float, out):
        a = np.ascontiguousarray(frames_bgr, dtype=np.uint8)
        single = a.ndim == 3
        if single:
            a = a[None]
        if a.shape != (self.n, self.h, self.w, 3):
            raise ValueError(f"frames must be uint8 [{self.n},{self.h},{self.w},3], got {a.shape}")
        if out is not None:
            if out.dtype != np.uint8 or out.shape != a.shape or not out.flags["C_CONTIGUOUS"]:
                raise ValueError("out must be a C-contiguous uint8 array of the frames' shape")
        else:
            out = np.empty_like(a)
        outf = np.empty(a.shape, np.float32) if return_float else None
        return a, out, outf, single

    @staticmethod
    def _result(out, outf, single):
        if single:
            out = out[0]
            outf = outf[0] if outf is not None else None
        return (out, outf) if outf is not None else out

    def step(self, frames_bgr, return_float=False, out=None):
        """frames_bgr: uint8 [H,W,3] (one clip) or [n_clips,H,W,3], BGR as cap.read() returns them.
        Returns np.uint8(totaloutputFrame[i]) with the same leading shape (and totaloutputFrame[i] as float32
        when return_float).  `out`: optional uint8 [n_clips,H,W,3] array to receive the frames (e.g. pinned_buffer())."""
        a, out, outf, single = self._check_args(frames_bgr, return_float, out)
        _lib.check(self._lib.ofs_clips_step_host(self._h, a.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                                                 outf.ctypes.data_as(C.c_void_p) if return_float else None))
        return self._result(out, outf, single)

    def submit(self, frames_bgr, return_float=False, out=None):
        """step() without the wait: queues the upload, the kernels and the download of one frame per clip and returns.
        At most `depth` (3) steps may be in flight; collect them in order with wait().  `frames_bgr` (and `out`) must
        not be written until the matching wait() returns -- use pinned_buffer() arrays, `depth` of each, in rotation."""
        a, out, outf, single = self._check_args(frames_bgr, return_float, out)
        _lib.check(self._lib.ofs_clips_submit_host(self._h, a.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                                                   outf.ctypes.data_as(C.c_void_p) if return_float else None))
        self._pending.append((a, out, outf, single))

    def submit_device(self, frames_u8, out_u8):
        """submit() for frames that already live on the GPU: `frames_u8` / `out_u8` are contiguous CUDA uint8 tensors
        [n_clips,H,W,3] (BGR); the stabilised frames land in `out_u8` on the device -- e.g. a slice of a per-rank
        [n_frames,H,W,3] buffer that sharding.gather_output() later collects over NCCL.  Same ordering rules as submit()."""
        import torch

        shape = (self.n, self.h, self.w, 3)
        for t, name in ((frames_u8, "frames_u8"), (out_u8, "out_u8")):
            if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.uint8 and t.is_contiguous() and tuple(t.shape) == shape):
                raise ValueError(f"{name} must be a contiguous CUDA uint8 tensor of shape {shape}")
        _lib.check(self._lib.ofs_clips_submit_host(self._h, C.c_void_p(frames_u8.data_ptr()), C.c_void_p(out_u8.data_ptr()), None))
        self._pending.append((frames_u8, out_u8, None, False))

    def wait(self):
        """Blocks until the oldest submitted step has landed and returns what step() would have returned for it."""
        if not self._pending:
            raise RuntimeError("ClipStabilizer.wait(): no step in flight")
        _lib.check(self._lib.ofs_clips_wait(self._h))
        _, out, outf, single = self._pending.popleft()
        return self._result(out, outf, single)
