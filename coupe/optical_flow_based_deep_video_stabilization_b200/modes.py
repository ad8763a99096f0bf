"""The reference's other test modes around the same network.

    evaluate_originalSize_homo()   main_flownetS_pyramid_noprevloss_dataloader.py:634-751   -> HomographyStabilizer
    evaluate()                     main_flownetS_pyramid_noprevloss_dataloader.py:758-866   -> FixedSizeStabilizer()
    evaluate_blurNma()             main_flownetS_pyramid.py:582-700                         -> FixedSizeStabilizer(flow_filter="blurNma")
    evaluate_medianNma()           main_flownetS_pyramid.py:703-820                         -> FixedSizeStabilizer(flow_filter="medianNma")

What the reference runs inside sess.run -- the network, the flow glue, tf_warp, the TF resize, the 75x75 mean filter --
and the per-frame OpenCV / SciPy array work on full frames (warpPerspective, the float resize, medfilt) run on the GPU
through libofstab.so.  What stays on the host, exactly as in the reference: decoding / encoding, the 512x384 uint8
cv2.resize + cvtColor that assemble the network input, and cv2.findHomography (RANSAC on the dense grid, main_dl.py:742).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .ops import _cuda_f32, tf_warp

STAB_OFFSETS_ORIGINAL = (31, 23, 15, 7, 4, 3, 2, 1)     # main_dl.py:707 (and :553)
STAB_OFFSETS_FIXED = (1, 2, 3, 4, 7, 15, 23, 31)         # main_dl.py:845, main_flownetS_pyramid.py:673


def _stream(dev):
    return _lib.current_stream_ptr(dev)


# ------------------------------------------------------------------------------------------- ops
def flow_resize_ex(flow, out_h, out_w, pre_mul):
    """resize_images(flow * pre_mul / fh, [out_h, out_w]) then x * out_w / 512, y * out_h / 384 (main_dl.py:497-498 with
    pre_mul = 384, :681-682 with flow = predict_flow3 and pre_mul = out_h)."""
    flow = _cuda_f32(flow, "flow")
    B, fh, fw, two = flow.shape
    if two != 2:
        raise ValueError("flow_resize_ex: last dim must be 2")
    out = torch.empty((B, int(out_h), int(out_w), 2), device=flow.device, dtype=torch.float32)
    with torch.cuda.device(flow.device):
        _lib.check(_lib.load().ofs_flow_resize_ex(_lib.ptr(flow), _lib.ptr(out), B, fh, fw, int(out_h), int(out_w),
                                                  float(pre_mul), _stream(flow.device)))
    return out


def warp_perspective_u8(frames, h, dsize=None):
    """cv2.warpPerspective(frame, h, (out_w, out_h)) for uint8 CUDA frames [B,H,W,3] (or [H,W,3]); h: [B,3,3] / [3,3]
    float64 host array (cv2.findHomography's result).  Byte-identical to OpenCV."""
    if not (isinstance(frames, torch.Tensor) and frames.is_cuda and frames.dtype == torch.uint8):
        raise TypeError("warp_perspective_u8: frames must be a CUDA uint8 tensor (no CPU fallback)")
    single = frames.dim() == 3
    fr = (frames[None] if single else frames).contiguous()
    B, H, W, Cc = fr.shape
    if Cc != 3:
        raise ValueError("warp_perspective_u8: 3-channel frames")
    hm = np.ascontiguousarray(np.asarray(h, dtype=np.float64).reshape(-1, 3, 3))
    if hm.shape[0] != B:
        raise ValueError(f"warp_perspective_u8: {hm.shape[0]} matrices for {B} frames")
    out_w, out_h = (W, H) if dsize is None else (int(dsize[0]), int(dsize[1]))
    out = torch.empty((B, out_h, out_w, 3), device=fr.device, dtype=torch.uint8)
    with torch.cuda.device(fr.device):
        _lib.check(_lib.load().ofs_warp_perspective_u8(C.c_void_p(fr.data_ptr()), hm.ctypes.data_as(C.c_void_p),
                                                       C.c_void_p(out.data_ptr()), B, H, W, out_h, out_w, _stream(fr.device)))
    return out[0] if single else out


def tf1_resize_images(x, size, c0=0, channels=None):
    """tf.image.resize_images(x[..., c0:c0+channels], size) -- TF-1.10 legacy bilinear (main_dl.py:806)."""
    x = _cuda_f32(x, "x")
    B, H, W, Ct = x.shape
    Cn = Ct - c0 if channels is None else int(channels)
    out = torch.empty((B, int(size[0]), int(size[1]), Cn), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().ofs_tf1_resize_bilinear(_lib.ptr(x), _lib.ptr(out), B, H, W, Ct, int(c0), Cn, int(size[0]),
                                                       int(size[1]), _stream(x.device)))
    return out


def cv_resize_linear(x, dsize, post_mul=1.0):
    """cv2.resize(float32 image, (w, h)) (INTER_LINEAR) * post_mul for CUDA float32 [B,H,W,C] (main_dl.py:862)."""
    x = _cuda_f32(x, "x")
    B, H, W, Cn = x.shape
    out = torch.empty((B, int(dsize[1]), int(dsize[0]), Cn), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().ofs_cv_resize_linear_f32(_lib.ptr(x), _lib.ptr(out), B, H, W, Cn, int(dsize[1]), int(dsize[0]),
                                                        float(post_mul), _stream(x.device)))
    return out


def flow_box_blur_ema(flow, prev=None, k=75, a=0.9, b=0.1):
    """a * conv2d(flow, 1/(k*k), SAME) + b * prev per flow plane (main_flownetS_pyramid.py:634-641); prev=None: the blur."""
    flow = _cuda_f32(flow, "flow")
    B, H, W, two = flow.shape
    if two != 2:
        raise ValueError("flow_box_blur_ema: flow must be [B,H,W,2]")
    if prev is not None:
        prev = _cuda_f32(prev, "prev")
        if prev.shape != flow.shape:
            raise ValueError("flow_box_blur_ema: prev must have the flow's shape")
    out, scratch = torch.empty_like(flow), torch.empty_like(flow)
    with torch.cuda.device(flow.device):
        _lib.check(_lib.load().ofs_flow_box_blur_ema(_lib.ptr(flow), _lib.ptr(prev), _lib.ptr(out), _lib.ptr(scratch), B, H, W,
                                                     int(k), float(a), float(b), _stream(flow.device)))
    return out


def medfilt(vol, kernel_size=5):
    """scipy.signal.medfilt(vol, kernel_size) for a CUDA float32 [H,W,C] array and a scalar odd kernel size."""
    vol = _cuda_f32(vol, "vol", ndim=3)
    H, W, Cn = vol.shape
    out = torch.empty_like(vol)
    with torch.cuda.device(vol.device):
        _lib.check(_lib.load().ofs_medfilt_nd3(_lib.ptr(vol), _lib.ptr(out), H, W, Cn, int(kernel_size), _stream(vol.device)))
    return out


# ------------------------------------------------------------------------------------------- per-frame drivers
def _resize_rgbswap_u8(img_u8):
    """cv2.cvtColor(cv2.resize(img, (512, 384)), COLOR_RGB2BGR) -- host, as the reference does it (main_dl.py:705,712)."""
    import cv2

    return cv2.cvtColor(cv2.resize(img_u8, (512, 384)), cv2.COLOR_RGB2BGR)


class HomographyStabilizer:
    """One clip through evaluate_originalSize_homo() (main_dl.py:696-748).  step(frame_bgr_u8) returns
    (np.uint8(curframeHomo), h): the frame written to the video and the fitted homography."""

    def __init__(self, net, height, width):
        self.net, self.h, self.w = net, int(height), int(width)
        self.history = []            # totaloutputFrame: float64 [H,W,3] per frame, as the reference keeps it (:694)
        self.dev = net.device

    def step(self, frame_unstab):
        import cv2

        i = len(self.history)
        out_h, out_w = self.h, self.w
        frame_unstab = np.ascontiguousarray(frame_unstab, dtype=np.uint8)
        if frame_unstab.shape != (out_h, out_w, 3):
            raise ValueError(f"frame must be uint8 [{out_h},{out_w},3]")
        first = frame_unstab.astype(np.float64) if i == 0 else self.history[0]                      # :703-704
        curinput = np.zeros([1, 384, 512, 27])
        curinput[0, :, :, 24:27] = _resize_rgbswap_u8(frame_unstab) / 255.0                         # :705
        for j, off in enumerate(STAB_OFFSETS_ORIGINAL):                                             # :707-712
            src = first if i - off < 0 else self.history[i - off]
            curinput[0, :, :, j * 3:(j + 1) * 3] = np.float32(_resize_rgbswap_u8(np.uint8(src))) / 255.0
        feats = torch.from_numpy(curinput.astype(np.float32)).to(self.dev)
        frame_dev = torch.from_numpy(frame_unstab).to(self.dev)
        resized_input = torch.from_numpy((cv2.cvtColor(frame_unstab, cv2.COLOR_RGB2BGR) / 255.0).astype(np.float32))[None].to(self.dev)
        flows = self.net.forward(feats)
        outflow = flow_resize_ex(flows["predict_flow3"], out_h, out_w, pre_mul=float(out_h))        # :681-682
        warped = tf_warp(resized_input, outflow, out_h, out_w)                                      # :683, sess.run :715
        curoutflow = outflow[0].cpu().numpy()                                                       # :714,716
        xv, yv = np.meshgrid(np.linspace(0, out_w - 1, out_w), np.linspace(0, out_h - 1, out_h))    # :719-721
        gridmesh = np.concatenate((np.expand_dims(xv, 2), np.expand_dims(yv, 2)), axis=2)
        gridmesh_of = gridmesh - curoutflow
        h, _ = cv2.findHomography(np.reshape(gridmesh, (out_w * out_h, 2)), np.reshape(gridmesh_of, (out_w * out_h, 2)),
                                  cv2.RANSAC)                                                       # :742 (host, as in the reference)
        cur_homo = warp_perspective_u8(frame_dev, h, (out_w, out_h))                                # :743 on the GPU
        self.last = {"predict_flow3": flows["predict_flow3"], "outflow": outflow, "warped": warped}  # for inspection / parity tests
        total = cv2.cvtColor(np.squeeze(warped.cpu().numpy()) * 255, cv2.COLOR_RGB2BGR)             # :747
        if i == 0:
            self.history.append(first)
            self.history[0] = total.astype(np.float64)      # :704 seeds slot 0, :747 then overwrites it with frame 0's result
        else:
            self.history.append(total.astype(np.float64))
        return np.uint8(cur_homo.cpu().numpy()), h                                                  # :751


class FixedSizeStabilizer:
    """One clip through evaluate() (main_dl.py:833-864) or, with flow_filter, evaluate_blurNma / evaluate_medianNma of the
    sibling driver (main_flownetS_pyramid.py:660-700 / 782-820).  Everything happens at the network's 512x384.
    step(frame_bgr_u8) returns the uint8 side-by-side frame [384, 1024, 3] the reference writes."""

    def __init__(self, net, flow_filter=None):
        if flow_filter not in (None, "blurNma", "medianNma"):
            raise ValueError("flow_filter: None, 'blurNma' or 'medianNma'")
        self.net, self.flow_filter, self.dev = net, flow_filter, net.device
        self.history = []            # totaloutputFrame at 384x512 (float64), :829
        self.prevof = torch.zeros((1, 382, 510, 2), device=self.dev)                                # main_flownetS_pyramid.py:658

    def step(self, frame_unstab):
        import cv2

        i = len(self.history)
        small = cv2.resize(np.ascontiguousarray(frame_unstab, dtype=np.uint8), (512, 384))
        first = small.astype(np.float64) if i == 0 else self.history[0]                             # :840-841
        curinput = np.zeros([1, 384, 512, 27])
        curinput[0, :, :, 24:27] = cv2.cvtColor(small, cv2.COLOR_RGB2BGR) / 255.0                   # :842
        for j, off in enumerate(STAB_OFFSETS_FIXED):                                                # :845-850
            src = first if i - off < 0 else self.history[i - off]
            curinput[0, :, :, j * 3:(j + 1) * 3] = np.float32(cv2.cvtColor(np.uint8(src), cv2.COLOR_RGB2BGR)) / 255.0
        feats = torch.from_numpy(curinput.astype(np.float32)).to(self.dev)
        of = self.net.forward(feats)["predict_flow2"]
        unstabimg = tf1_resize_images(feats, (382, 510), c0=24, channels=3)                         # :806
        if self.flow_filter is None:
            flow = of                                                                               # :807
        elif self.flow_filter == "blurNma":
            flow = flow_box_blur_ema(of, self.prevof, k=75, a=0.9, b=0.1)                           # main_flownetS_pyramid.py:634-641
            self.prevof = 0.9 * self.prevof + 0.1 * of                                              # :694
        else:
            med = medfilt(of[0], 5)[None]                                                           # :809
            flow = 0.9 * med + 0.1 * self.prevof                                                    # :759
            self.prevof = 0.9 * self.prevof + 0.1 * med                                             # :813
        flow = flow.contiguous()
        warped = tf_warp(unstabimg, flow, 382, 510)
        self.last = {"predict_flow2": of, "flow": flow, "unstabimg": unstabimg, "warped": warped}   # for inspection / parity tests
        total = cv_resize_linear(warped, (512, 384), post_mul=255.0)[0].cpu().numpy()               # :862 cv2.resize(...) * 255
        total = cv2.cvtColor(total, cv2.COLOR_RGB2BGR)
        if i == 0:
            self.history.append(first)
            self.history[0] = total.astype(np.float64)
        else:
            self.history.append(total.astype(np.float64))
        return np.uint8(np.concatenate([np.uint8(small), np.uint8(total)], axis=1))                 # :864
