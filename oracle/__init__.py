"""CPU oracle for the stabilizer's inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / CPU baseline.
The product path (``coupe.optical_flow_based_deep_video_stabilization_b200``)
never imports this package and fails loudly when its CUDA library is missing.

PARITY UNPINNED.  The reference (TensorFlow 1.10 + TensorLayer 1.x, pure
Python) cannot be imported or executed in this image (no ``tensorflow`` /
``tensorlayer`` wheels, no network) and it ships no tests, golden vectors or
fixtures for this path.  The oracle is therefore a restatement of the
reference's arithmetic, written from the cited reference lines and from the
published TF-1.10 / TensorLayer-1.x op semantics:

  * ``oracle.tf1_ops``   vectorised torch restatement of the third-party ops
                         the path calls (legacy bilinear resize, NN resize with
                         align_corners, conv2d VALID on explicit zero pad,
                         conv2d_transpose SAME k4 s2, BN inference, lrelu).
  * ``oracle.literal``   an independent, slow, loop-level restatement of the
                         same ops in pure numpy, used to cross-check
                         ``tf1_ops``/``samplers`` on small inputs (two separately
                         written restatements must agree).
  * ``oracle.flownet``   ``flownetS_pyramid``  (reference ``model.py:786-893``)
  * ``oracle.samplers``  ``tf_warp`` (``main_flownetS_pyramid_noprevloss_dataloader.py:44-130``),
                         test-mode flow glue (``...dataloader.py:497-514``),
                         ``bilinear_interp`` / ``AffineTransformer`` /
                         ``ProjectiveTransformer`` (``spatial_transformer.py:373-452,
                         519-608,755-779,902-964``), ``vec2mtrx`` /
                         ``transformImage`` / ``transformCropImage``
                         (``warp.py:25-129``).

Known-answer vectors derived analytically from those lines live in
``tests/golden/`` together with the script that generates them.
"""
