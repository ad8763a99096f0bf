"""Host mirror of the reference's model entry point.

    flownetS_pyramid(feats, batch_size, is_train=False, reuse=False, scope='flownetS') -> dict
                                                                        reference model.py:786-893

The reference builds a TF graph whose variables are later filled by
``tl.files.load_and_assign_npz_dict`` (main_flownetS_pyramid_noprevloss_dataloader.py:520).
Here the call is eager: weights are assigned to a *scope* first (``load_and_assign_npz_dict`` /
``assign_weights``), then ``flownetS_pyramid`` runs the CUDA forward and returns the same 6-key
dict of NHWC float32 CUDA tensors.  ``FlowNetSPyramid`` is the underlying handle
(one per GPU and stream, not re-entrant).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .ops import _cuda_f32

NET_H, NET_W, NET_C = 384, 512, 27
FLOW_SHAPES = {6: (6, 8), 5: (12, 16), 4: (24, 32), 3: (48, 64), 2: (382, 510)}

_ACT_NAMES = ["input", "conv1", "conv2", "conv3", "conv3_1", "conv4", "conv4_1", "conv5", "conv5_1", "conv6",
              "conv6_1", "concat5", "concat4", "concat3", "concat2"]


class FlowNetSPyramid:
    """Device-resident FlowNetS-pyramid (packed weights, activation workspace, TMA descriptors)."""

    def __init__(self, device=None, max_batch=8, precision="bf16"):
        if not torch.cuda.is_available():
            raise RuntimeError("FlowNetSPyramid needs a CUDA device (sm_100a); there is no CPU fallback")
        if isinstance(device, str):
            device = torch.device(device)
        if isinstance(device, torch.device):
            device = device.index                       # torch.device("cuda") has index None = the current device
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        self.max_batch = int(max_batch)
        self.precision = precision
        self._lib = _lib.load()
        self._h = C.c_void_p()
        prec = {"bf16": _lib.PREC_BF16, "fp16": _lib.PREC_FP16}[precision]
        _lib.check(self._lib.ofs_net_create(C.byref(self._h), self.device.index, self.max_batch, prec))
        self.loaded = False
        self._weights = None     # the mapping last assigned (by reference): lets get_net() rebuild a scope at a larger batch
        self._users = 0          # ClipStabilizers holding the raw ofs_net* of this handle (see retain / release)

    def retain(self):
        self._users += 1

    def release(self):
        self._users = max(0, self._users - 1)

    def close(self):
        if getattr(self, "_users", 0):
            raise RuntimeError(f"FlowNetSPyramid.close: {self._users} ClipStabilizer(s) still hold this net; close them first")
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.ofs_net_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self._users = 0
            self.close()
        except Exception:
            pass

    # ---------------------------------------------------------------- checkpoint ingest
    def assign_weights(self, weights):
        """weights: mapping TF-variable-name -> float32 array (TensorLayer npz dict layout)."""
        keep, arr = [], (_lib.NamedArray * len(weights))()
        for i, (name, val) in enumerate(weights.items()):
            a = np.ascontiguousarray(np.asarray(val, dtype=np.float32))
            keep.append(a)
            arr[i].name = str(name).encode()
            arr[i].data = a.ctypes.data_as(C.c_void_p)
            arr[i].numel = a.size
        self.loaded = False
        _lib.check(self._lib.ofs_net_load_weights(self._h, arr, len(weights)))
        self.loaded = True
        self._weights = weights

    def load_npz(self, path):
        with np.load(path, allow_pickle=False) as z:
            self.assign_weights({k: z[k] for k in z.files})

    # ---------------------------------------------------------------------- execution
    def forward(self, feats):
        feats = _cuda_f32(feats, "feats")
        B = feats.shape[0]
        if tuple(feats.shape[1:]) != (NET_H, NET_W, NET_C):
            raise ValueError(f"feats must be [B,{NET_H},{NET_W},{NET_C}] (model.py:850 hard-wires the pyramid), "
                             f"got {tuple(feats.shape)}")
        outs = {l: torch.empty((B, h, w, 2), device=feats.device, dtype=torch.float32)
                for l, (h, w) in FLOW_SHAPES.items()}
        with torch.cuda.device(feats.device):
            _lib.check(self._lib.ofs_net_forward(self._h, _lib.ptr(feats), B, _lib.ptr(outs[6]), _lib.ptr(outs[5]),
                                                 _lib.ptr(outs[4]), _lib.ptr(outs[3]), _lib.ptr(outs[2]),
                                                 _lib.current_stream_ptr(feats.device)))
        return {"predict_flow6": outs[6], "predict_flow5": outs[5], "predict_flow4": outs[4],
                "predict_flow3": outs[3], "predict_flow2": outs[2], "flow": outs[2]}

    def stabilize(self, feats, frames, return_flow=False):
        """One sess.run(outputs_warpedimg) of the test loop: forward + flow glue + warp on device tensors."""
        feats = _cuda_f32(feats, "feats")
        frames = _cuda_f32(frames, "frames")
        B, H, W, Cc = frames.shape
        if Cc != 3 or feats.shape[0] != B:
            raise ValueError("stabilize: frames must be [B,H,W,3] with the batch of feats")
        out = torch.empty_like(frames)
        flow2 = torch.empty((B, 382, 510, 2), device=feats.device, dtype=torch.float32) if return_flow else None
        with torch.cuda.device(feats.device):
            _lib.check(self._lib.ofs_net_stabilize(self._h, _lib.ptr(feats), _lib.ptr(frames), _lib.ptr(out),
                                                   _lib.ptr(flow2), B, H, W, _lib.current_stream_ptr(feats.device)))
        return (out, flow2) if return_flow else out

    def stabilize_host(self, feats, frames, out=None):
        """Same call on HOST float32 tensors/arrays (pinned memory recommended): H2D, compute, D2H."""
        feats = _host_f32(feats, "feats")
        frames = _host_f32(frames, "frames")
        B, H, W, Cc = frames.shape
        if Cc != 3 or feats.shape[0] != B or tuple(feats.shape[1:]) != (NET_H, NET_W, NET_C):
            raise ValueError("stabilize_host: bad shapes")
        if out is None:
            out = torch.empty_like(frames)
        _lib.check(self._lib.ofs_net_stabilize_host(self._h, C.c_void_p(feats.data_ptr()),
                                                    C.c_void_p(frames.data_ptr()), C.c_void_p(out.data_ptr()),
                                                    B, H, W))
        return out

    def activation(self, name, batch):
        """Parity/debug: a named intermediate of the LAST forward, widened to float32 [B,H,W,C]."""
        shape = (C.c_int * 4)()
        cap = batch * 384 * 512 * 32
        buf = torch.empty(cap, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.ofs_net_get_activation(self._h, name.encode(), batch, _lib.ptr(buf), cap, shape,
                                                        _lib.current_stream_ptr(self.device)))
        n = shape[0] * shape[1] * shape[2] * shape[3]
        return buf[:n].view(*list(shape)).clone()

    def profile(self, feats, frames=None, iters=5):
        """Mean device time of every kernel launch of one step: list of (name, ms, macs)."""
        feats = _cuda_f32(feats, "feats")
        B = feats.shape[0]
        H = W = 0
        out = None
        if frames is not None:
            frames = _cuda_f32(frames, "frames")
            _, H, W, _ = frames.shape
            out = torch.empty_like(frames)
        cap = 64
        ms = (C.c_float * cap)()
        macs = (C.c_double * cap)()
        names = C.create_string_buffer(cap * 32)
        count = C.c_int(0)
        with torch.cuda.device(feats.device):
            _lib.check(self._lib.ofs_net_profile(self._h, _lib.ptr(feats), _lib.ptr(frames), _lib.ptr(out), B, H, W,
                                                 int(iters), C.cast(ms, C.c_void_p), C.cast(macs, C.c_void_p),
                                                 C.cast(names, C.c_void_p), cap, C.byref(count),
                                                 _lib.current_stream_ptr(feats.device)))
        res = []
        for i in range(count.value):
            nm = names.raw[i * 32:(i + 1) * 32].split(b"\0", 1)[0].decode()
            res.append((nm, float(ms[i]), float(macs[i])))
        return res

    def time_kernels(self, which, batch, frames=None, iters=20):
        """Device time per repetition of a kernel set replayed from one CUDA graph (see ofs_net_time_kernels):
        which = "dense" -> (ms, macs, launches) of the 14 conv / deconv GEMM launches; "warp" -> fused warp."""
        ms = C.c_float(0)
        macs = C.c_double(0)
        nl = C.c_int(0)
        H = W = 0
        out = None
        if which == "warp":
            frames = _cuda_f32(frames, "frames")
            _, H, W, _ = frames.shape
            out = torch.empty_like(frames)
        dev = torch.device("cuda", self.device) if not isinstance(self.device, torch.device) else self.device
        with torch.cuda.device(dev):
            _lib.check(self._lib.ofs_net_time_kernels(self._h, 0 if which == "dense" else 1, _lib.ptr(frames), _lib.ptr(out),
                                                      int(batch), H, W, int(iters), C.byref(ms), C.byref(macs), C.byref(nl)))
        return float(ms.value), float(macs.value), int(nl.value)

    def graph_stats(self):
        """(captures, re-pointings, cached graphs) of the stabilize() step-graph cache."""
        return tuple(int(self._lib.ofs_net_graph_stats(self._h, k)) for k in range(3))

    @property
    def launches_per_forward(self):
        return int(self._lib.ofs_net_launches_per_forward(self._h))


def _host_f32(t, name):
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(t, dtype=np.float32))
    if t.is_cuda or t.dtype != torch.float32:
        raise TypeError(f"{name}: expected a host float32 tensor")
    return t.contiguous()


# ---------------------------------------------------------------------------------------------
# reference-shaped functional API (scope registry stands in for TF variable scopes)
_NETS = {}


def _scope_key(scope, device):
    idx = None
    if device is not None:
        idx = device if isinstance(device, int) else torch.device(device).index
    return (scope, torch.cuda.current_device() if idx is None else int(idx))   # "cuda" (index None) = the current device


def get_net(scope="flownetS", device=None, max_batch=8, precision="bf16"):
    """The net of a scope (created on first use).  Asking for a larger max_batch or another precision REBUILDS the
    scope's net: the weights assigned to the old one are re-assigned to the new one, and the old handle is destroyed
    -- which is refused while a ClipStabilizer still holds it (its C side keeps the raw ofs_net*)."""
    key = _scope_key(scope, device)
    net = _NETS.get(key)
    if net is None or net.max_batch < max_batch or net.precision != precision:
        old = net
        if old is not None and old._users:
            raise RuntimeError(f"scope '{scope}': cannot rebuild the net (max_batch {old.max_batch} -> {max_batch}, precision "
                               f"{old.precision} -> {precision}) while {old._users} ClipStabilizer(s) use it; close them first")
        net = FlowNetSPyramid(device=key[1], max_batch=max(max_batch, old.max_batch if old else 1), precision=precision)
        if old is not None:
            if old.loaded and old._weights is not None:
                net.assign_weights(old._weights)
            old.close()
        _NETS[key] = net
    return net


def assign_weights(weights, scope="flownetS", device=None, max_batch=8, precision="bf16"):
    get_net(scope, device, max_batch, precision).assign_weights(weights)


def load_and_assign_npz_dict(name, sess=None, scope="flownetS", device=None, max_batch=8, precision="bf16"):
    """tl.files.load_and_assign_npz_dict(name=..., sess=...) of the reference (main:520); `sess` is ignored."""
    get_net(scope, device, max_batch, precision).load_npz(name)


def flownetS_pyramid(feats, batch_size, is_train=False, reuse=False, scope="flownetS"):
    """Reference signature (model.py:786).  Returns {'predict_flow6', ..., 'predict_flow2', 'flow'}.

    ``is_train`` must be False (inference tier: BatchNorm uses moving statistics); ``reuse`` is accepted
    and ignored; ``batch_size`` must equal feats.shape[0] (the reference bakes it into deconv shapes).
    """
    if is_train:
        raise NotImplementedError("flownetS_pyramid: only is_train=False (inference) is implemented")
    if int(batch_size) != int(feats.shape[0]):
        raise ValueError(f"batch_size {batch_size} != feats.shape[0] {feats.shape[0]}")
    key = _scope_key(scope, feats.device)
    net = _NETS.get(key)
    if net is None or not net.loaded:
        raise RuntimeError(f"no weights assigned to scope '{scope}': call load_and_assign_npz_dict / assign_weights first")
    if net.max_batch < feats.shape[0]:
        raise RuntimeError(f"scope '{scope}' was created with max_batch {net.max_batch} < {feats.shape[0]}")
    return net.forward(feats)
