"""B200-native SSD multibox head path (match + mined CE/L1 loss fwd/bwd, decode + NMS).

A from-scratch sm_100a implementation of the hot path of nitishsaDire/objectDetection_ssd
(``Losses.py`` / ``Util.py``), reached through the C ABI in ``include/ssdhead.h``.
``objectdetection_ssd_b200.Losses`` and ``objectdetection_ssd_b200.Util`` mirror the
reference's call surface; ``objectdetection_ssd_b200/dropin`` makes them importable under
the reference's top-level module names.
"""
__version__ = "0.1.0"
