"""Video boundary of the test mode (reference main_flownetS_pyramid_noprevloss_dataloader.py:451-632,
`evaluate_originalSize()`): cv2.VideoCapture in, MJPG AVI out, the loop body on the GPU (ClipStabilizer).

    stabilize_video("0.avi", "result_video/0_out.avi", ckpt="fixed_ckpt/flownetS_pyramid.npz")

Kept from the reference: total_frames = CAP_PROP_FRAME_COUNT - 2 (:479), the writer (fourcc MJPG, the source fps and
size, :483-487), one np.uint8 frame written per input frame (:630).  Not kept: the hard-coded list of six test
videos (:472) and the absolute cluster paths of config.py -- the caller names the files.
"""
from __future__ import annotations

import os

from .clip import ClipStabilizer
from .model import get_net, load_and_assign_npz_dict


def stabilize_video(src_path, dst_path, ckpt=None, net=None, scope="flownetS", device=None, max_frames=None):
    """Stabilises one video file.  `ckpt`: TensorLayer npz (tl.files.save_npz_dict format) loaded into `scope`
    unless `net` (a FlowNetSPyramid with weights) is given.  Returns the number of frames written."""
    import cv2

    if net is None:
        if ckpt is not None:
            load_and_assign_npz_dict(name=ckpt, sess=None, scope=scope, device=device, max_batch=1)
        net = get_net(scope=scope, device=device, max_batch=1)
    cap = cv2.VideoCapture(src_path)                                                 # main_dl.py:477
    if not cap.isOpened():
        raise IOError(f"cannot open video {src_path!r}")
    fps = cap.get(5)                                                                 # :478
    total_frames = int(cap.get(7) - 2)                                               # :479
    out_h, out_w = int(cap.get(4)), int(cap.get(3))                                  # :480-481
    if max_frames is not None:
        total_frames = min(total_frames, int(max_frames))
    os.makedirs(os.path.dirname(os.path.abspath(dst_path)) or ".", exist_ok=True)
    out = cv2.VideoWriter(dst_path, cv2.VideoWriter_fourcc("M", "J", "P", "G"), fps, (out_w, out_h))   # :483-487
    stab = ClipStabilizer(net, n_clips=1, height=out_h, width=out_w)
    # page-locked frame / result buffers in rotation: frame i+1 is decoded and uploaded, frame i-1 downloaded and
    # encoded, while the GPU works on frame i
    depth = stab.depth
    fin = [stab.pinned_buffer() for _ in range(depth)]
    fout = [stab.pinned_buffer() for _ in range(depth)]
    written = 0
    try:
        for i in range(max(total_frames, 0)):                                        # main_dl.py:540
            ret_unstab, frame_unstab = cap.read()                                    # :547
            if not ret_unstab:
                break                                                                # the reference would crash in cv2.resize
            if stab.in_flight == depth:
                out.write(stab.wait()[0])                                            # :630
                written += 1
            fin[i % depth][0] = frame_unstab
            stab.submit(fin[i % depth], out=fout[i % depth])                                 # :550-625
        while stab.in_flight:
            out.write(stab.wait()[0])
            written += 1
    finally:
        out.release()                                                                # :632
        cap.release()
        stab.close()
    return written
