"""Top-level ``Util`` shim (see ``dropin/Losses.py``).  Names outside the head path (dataset lists,
``transform``, drawing, ``get_map``) forward to the reference's own Util when SSD_REFERENCE_DIR points at it."""
from objectdetection_ssd_b200 import Util as _impl
from objectdetection_ssd_b200.Util import *            # noqa: F401,F403
from objectdetection_ssd_b200.Util import (device, class_to_label, label_to_class, create_priors_ssd300,  # noqa: F401
                                           xywh_to_xyxy, xyxy_to_xywh, gcxgcy_to_cxcy, get_offsets_coords,
                                           find_intersection, get_jaccard_tensor1, get_jaccard_tensor11,
                                           map_prior_to_bb, subsampling)


def __getattr__(name):
    return getattr(_impl, name)
