"""poco_b200 -- B200-native (sm_100a) implementation of POCO's per-crop inference hot path.

    from poco_b200 import POCO        # drop-in for `from pocolib.models import POCO`
"""
from ._lib import PocoError, kernel_launches  # noqa: F401
from .poco import POCO  # noqa: F401
from .preprocess import crop_batch, uncert_post  # noqa: F401
from .smpl import DeviceSmplStage, load_smpl_model  # noqa: F401
from .stream import StreamRunner, convert_crop_cam_to_orig_img  # noqa: F401

__all__ = ['POCO', 'PocoError', 'kernel_launches', 'crop_batch', 'uncert_post', 'DeviceSmplStage', 'load_smpl_model', 'StreamRunner',
           'convert_crop_cam_to_orig_img']
