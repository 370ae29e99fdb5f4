"""Top-level ``Losses`` shim: put ``objectdetection_ssd_b200/dropin`` first on PYTHONPATH and the reference's
``train_function.py`` (``from Losses import *``) picks up the B200 implementation unchanged."""
from objectdetection_ssd_b200 import Losses as _impl
from objectdetection_ssd_b200.Losses import *          # noqa: F401,F403
from objectdetection_ssd_b200.Losses import (ancs_xywh, ancs_xyxy, device, ssd, ssd1_, ssd_old, ssd1, inference,  # noqa: F401
                                             inference_batch)


def __getattr__(name):
    return getattr(_impl, name)
