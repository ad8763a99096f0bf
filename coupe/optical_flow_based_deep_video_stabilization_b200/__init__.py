"""B200-native (sm_100a) inference hot path of optical-flow deep video stabilisation.

Drop-in host mirror of the reference call signatures, backed by libofstab.so (C ABI, include/ofstab.h):

    from coupe.optical_flow_based_deep_video_stabilization_b200 import (
        flownetS_pyramid, load_and_assign_npz_dict, tf_warp, AffineTransformer, transformImage)

No CPU fallback, no Triton, no backend dispatch: importing is cheap, the first device call loads
(and if needed builds) the CUDA library and raises if that is impossible.
"""
from ._lib import OfstabError, lib_path, load as load_library            # noqa: F401
from .clip import ClipStabilizer                                                               # noqa: F401
from .driver import stabilize_video                                                            # noqa: F401
from .model import (FlowNetSPyramid, assign_weights, flownetS_pyramid, get_net,                 # noqa: F401
                    load_and_assign_npz_dict)
from .modes import (FixedSizeStabilizer, HomographyStabilizer, cv_resize_linear, flow_box_blur_ema,  # noqa: F401
                    flow_resize_ex, medfilt, tf1_resize_images, warp_perspective_u8)
from .ops import (conv2d_nhwc, flow_resize, flow_resize_warp, get_pixel_value, set_warp_variant,  # noqa: F401
                  tf_warp)
from .sharding import gather_output, shard_range                                                # noqa: F401
from .spatial_transformer import AffineTransformer, ProjectiveTransformer, transformer          # noqa: F401
from .warp import compose, fit, inverse, transformCropImage, transformImage, vec2mtrx           # noqa: F401

__all__ = [
    "FlowNetSPyramid", "flownetS_pyramid", "load_and_assign_npz_dict", "assign_weights", "get_net",
    "tf_warp", "get_pixel_value", "flow_resize", "flow_resize_warp", "set_warp_variant", "conv2d_nhwc",
    "AffineTransformer", "ProjectiveTransformer", "transformer",
    "vec2mtrx", "transformImage", "transformCropImage", "fit", "compose", "inverse",
    "shard_range", "gather_output", "ClipStabilizer", "stabilize_video",
    "HomographyStabilizer", "FixedSizeStabilizer", "warp_perspective_u8", "flow_resize_ex", "tf1_resize_images",
    "cv_resize_linear", "flow_box_blur_ema", "medfilt",
    "OfstabError", "lib_path", "load_library",
]
