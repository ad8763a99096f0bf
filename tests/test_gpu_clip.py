"""Clip driver against a literal replay of the reference loop body,
main_flownetS_pyramid_noprevloss_dataloader.py:540-630, with cv2 / numpy on the host and sess.run replaced by the
(already parity-tested) fused GPU step.  The device-side driver must reproduce the written uint8 frames and the
float32 history byte for byte: same cv2.resize fixed-point arithmetic, same /255 variants, same np.uint8 cast."""
import numpy as np
import pytest
import torch

from oracle import flownet as F

cv2 = pytest.importorskip("cv2")
pytestmark = pytest.mark.gpu

STABIDXS = [31, 23, 15, 7, 4, 3, 2, 1]                                              # main_dl.py:553


@pytest.fixture(scope="module")
def ofs(cuda_dev):
    import coupe.optical_flow_based_deep_video_stabilization_b200 as m

    m.load_library()
    return m


def reference_loop(net, frames, dev):
    """frames: uint8 [T,H,W,3] BGR (what cap.read() returns).  Returns (written uint8 [T,H,W,3], totaloutputFrame)."""
    T, out_h, out_w, _ = frames.shape
    total = np.zeros([T, out_h, out_w, 3])                                           # :535 (float64)
    written = np.zeros([T, out_h, out_w, 3], np.uint8)
    for i in range(T):                                                               # :540
        curinput = np.zeros([1, 384, 512, 27])                                       # :544
        frame_unstab = frames[i]                                                     # :547
        if i == 0:
            total[0] = frame_unstab                                                  # :548-549
        curinput[0, :, :, 24:27] = cv2.cvtColor(cv2.resize(frame_unstab, (512, 384)), cv2.COLOR_RGB2BGR) / 255.0   # :550
        for j in range(len(STABIDXS)):                                               # :554-558
            idx = 0 if i - STABIDXS[j] < 0 else i - STABIDXS[j]
            with np.errstate(invalid="ignore"):
                hist = np.uint8(total[idx])
            curinput[0, :, :, j * 3:(j + 1) * 3] = np.float32(cv2.cvtColor(cv2.resize(hist, (512, 384)), cv2.COLOR_RGB2BGR)) / 255.0
        resized = np.expand_dims(cv2.cvtColor(frame_unstab, cv2.COLOR_RGB2BGR) / 255.0, axis=0)                     # :568
        feats = torch.from_numpy(curinput.astype(np.float32)).to(dev)                # feed_dict -> float32 placeholders
        frame = torch.from_numpy(resized.astype(np.float32)).to(dev)
        curwarpedimg = net.stabilize(feats, frame).cpu().numpy()                     # :569 sess.run(outputs_warpedimg)
        total[i] = cv2.cvtColor(np.squeeze(curwarpedimg) * 255, cv2.COLOR_RGB2BGR)   # :625
        with np.errstate(invalid="ignore"):
            written[i] = np.uint8(total[i])                                          # :630 out.write(...)
    return written, total


def synth_clip(seed, T, H, W):
    """A smooth, slowly drifting scene with frame-to-frame jitter: values stay well inside [0, 255]."""
    rng = np.random.default_rng(seed)
    base = cv2.GaussianBlur(rng.integers(0, 256, (H + 32, W + 32, 3)).astype(np.float32), (0, 0), 3.0)
    base = (base - base.min()) / (base.max() - base.min()) * 200 + 25
    out = np.zeros((T, H, W, 3), np.uint8)
    for i in range(T):
        dy, dx = rng.integers(0, 9, 2) + 8
        out[i] = base[dy:dy + H, dx:dx + W].astype(np.uint8)
    return out


@pytest.mark.parametrize("n_clips,T,H,W", [(1, 36, 96, 128), (2, 8, 120, 160), (1, 5, 90, 126)],
                         ids=["one_clip_ring_wraps", "two_clips_lockstep", "width_not_a_multiple_of_4"])
def test_clip_driver_matches_reference_loop(ofs, cuda_dev, n_clips, T, H, W):
    w = F.make_weights(0, "calibrated", head_scale=0.02)
    net = ofs.FlowNetSPyramid(device=cuda_dev, max_batch=max(2, n_clips))
    net.assign_weights(w)
    clips = [synth_clip(100 + c, T, H, W) for c in range(n_clips)]
    want = [reference_loop(net, clip, cuda_dev) for clip in clips]
    stab = ofs.ClipStabilizer(net, n_clips=n_clips, height=H, width=W)
    for i in range(T):
        frames = np.stack([clip[i] for clip in clips])
        got_u8, got_f32 = stab.step(frames, return_float=True)
        assert stab.frame_index == i + 1
        for c in range(n_clips):
            np.testing.assert_array_equal(got_u8[c], want[c][0][i], err_msg=f"clip {c} frame {i}: written frame")
            np.testing.assert_array_equal(got_f32[c], want[c][1][i].astype(np.float32), err_msg=f"clip {c} frame {i}: history")
    # reset starts the clip over and reproduces frame 0
    stab.reset()
    again = stab.step(np.stack([clip[0] for clip in clips]))
    for c in range(n_clips):
        np.testing.assert_array_equal(again[c], want[c][0][0])
    with pytest.raises(ValueError):
        stab.step(np.zeros((n_clips, H + 1, W, 3), np.uint8))
    stab.close()
    net.close()


def test_pipelined_submit_wait_equals_stepping(ofs, cuda_dev):
    """submit / wait (`depth` steps in flight, upload and download overlapping the kernels) returns exactly what the
    synchronous step returns, frame for frame, through a ring wrap; depth and ordering errors are reported."""
    n_clips, T, H, W = 2, 40, 96, 128
    net = ofs.FlowNetSPyramid(device=cuda_dev, max_batch=2)
    net.assign_weights(F.make_weights(0, "calibrated", head_scale=0.02))
    clips = np.stack([synth_clip(300 + c, T, H, W) for c in range(n_clips)])         # [n,T,H,W,3]
    sync = ofs.ClipStabilizer(net, n_clips=n_clips, height=H, width=W)
    want = [sync.step(np.ascontiguousarray(clips[:, i]), return_float=True) for i in range(T)]
    sync.close()
    pipe = ofs.ClipStabilizer(net, n_clips=n_clips, height=H, width=W)
    depth = pipe.depth
    assert depth >= 2
    fin = [pipe.pinned_buffer() for _ in range(depth)]
    fout = [pipe.pinned_buffer() for _ in range(depth)]
    got = []
    for i in range(T):
        if pipe.in_flight == depth:
            u8, f32 = pipe.wait()
            got.append((u8.copy(), f32.copy()))
        fin[i % depth][...] = clips[:, i]
        pipe.submit(fin[i % depth], return_float=True, out=fout[i % depth])
        assert pipe.frame_index == i + 1
    with pytest.raises(RuntimeError, match="in flight"):
        pipe.submit(fin[0], out=fout[0])                                             # one more than `depth` is refused
    with pytest.raises(RuntimeError, match="not yet waited"):
        pipe.step(fin[0])
    while pipe.in_flight:
        u8, f32 = pipe.wait()
        got.append((u8.copy(), f32.copy()))
    with pytest.raises(RuntimeError):
        pipe.wait()
    assert len(got) == T
    for i in range(T):
        np.testing.assert_array_equal(got[i][0], want[i][0], err_msg=f"frame {i}: written frame")
        np.testing.assert_array_equal(got[i][1], want[i][1], err_msg=f"frame {i}: history")
    # reset drains and restarts
    pipe.submit(fin[0], out=fout[0])
    pipe.reset()
    assert pipe.in_flight == 0 and pipe.frame_index == 0
    np.testing.assert_array_equal(pipe.step(np.ascontiguousarray(clips[:, 0])), want[0][0])
    pipe.close()
    net.close()


def test_every_byte_value_through_the_uint8_warp(ofs, cuda_dev):
    """Zero flow heads make the warp the identity in the interior, so the step's output is
    float32(float32(v / 255.0) * 255) per byte (main_dl.py:568, :625) -- for all 256 values v, which pins the
    on-the-fly v / 255 of the uint8 warp kernel to the reference's float64 quotient rounded to float32."""
    H, W = 96, 128
    net = ofs.FlowNetSPyramid(device=cuda_dev, max_batch=1)
    net.assign_weights(F.make_weights(0, "he", head_scale=0.0))
    frame = (np.arange(H * W * 3, dtype=np.int64) % 256).astype(np.uint8).reshape(H, W, 3)
    frame = np.ascontiguousarray(frame[:, ::-1])                                     # any arrangement holding all values
    assert len(np.unique(frame[:-1, :-1])) == 256
    stab = ofs.ClipStabilizer(net, n_clips=1, height=H, width=W)
    u8, f32 = stab.step(frame, return_float=True)
    want_f32 = (frame / 255.0).astype(np.float32) * np.float32(255)                  # cvtColor swaps cancel
    want_f32[-1, :] = 0                                                              # x1 == x0 / y1 == y0 at the far border:
    want_f32[:, -1] = 0                                                              # all four weights vanish (main_dl.py:96-118)
    np.testing.assert_array_equal(f32, want_f32)
    np.testing.assert_array_equal(u8, want_f32.astype(np.int32).astype(np.uint8))
    stab.close()
    net.close()


def test_stabilize_video_file_boundary(ofs, cuda_dev, tmp_path):
    """evaluate_originalSize()'s file handling (main_dl.py:477-487, :540-547, :630-632): MJPG AVI in, MJPG AVI out,
    CAP_PROP_FRAME_COUNT - 2 frames, each equal to what the clip driver returns for the decoded frames."""
    T, H, W = 12, 96, 128
    clip = synth_clip(7, T, H, W)
    src, dst = str(tmp_path / "in.avi"), str(tmp_path / "result_video" / "in_out.avi")
    wr = cv2.VideoWriter(src, cv2.VideoWriter_fourcc("M", "J", "P", "G"), 30.0, (W, H))
    assert wr.isOpened()
    for f in clip:
        wr.write(f)
    wr.release()
    w = F.make_weights(0, "calibrated", head_scale=0.02)
    net = ofs.FlowNetSPyramid(device=cuda_dev, max_batch=1)
    net.assign_weights(w)
    n = ofs.stabilize_video(src, dst, net=net)
    assert n == T - 2                                                                # :479
    # expected: the clip driver on the DECODED frames (MJPG is lossy), re-encoded by the same writer
    cap = cv2.VideoCapture(src)
    decoded = [cap.read()[1] for _ in range(T - 2)]
    cap.release()
    stab = ofs.ClipStabilizer(net, n_clips=1, height=H, width=W)
    want = [stab.step(f) for f in decoded]
    stab.close()
    ref_path = str(tmp_path / "ref.avi")
    wr = cv2.VideoWriter(ref_path, cv2.VideoWriter_fourcc("M", "J", "P", "G"), 30.0, (W, H))
    for f in want:
        wr.write(f)
    wr.release()
    a, b = cv2.VideoCapture(dst), cv2.VideoCapture(ref_path)
    assert int(a.get(7)) == T - 2 and int(a.get(3)) == W and int(a.get(4)) == H and abs(a.get(5) - 30.0) < 1e-6
    for _ in range(T - 2):
        ra, fa = a.read()
        rb, fb = b.read()
        assert ra and rb
        np.testing.assert_array_equal(fa, fb)
    a.release(); b.release()
    net.close()
