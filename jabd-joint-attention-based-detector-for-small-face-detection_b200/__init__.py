"""B200-native box-geometry hot path for the JABD small-face detector.

Import this package as ``jabd_b200`` (see ``/jabd_b200/__init__.py``); the
directory name mirrors the reference repository and is not an identifier.

Modules
-------
config              prior-box cfg dicts (contract of R/utils/config.py)
anchors             ``Anchors`` / ``Anchors_eval`` (R/utils/anchors.py)
retinaface_training ``match`` & friends, live training signatures
                    (R/nets/retinaface_training.py:8-162)
box_utils           SSD-legacy signatures (R/utils/box_utils.py)
utils_bbox          ``decode``/``decode_landm``/``non_max_suppression``/``nms_r``
                    (R/utils/utils_bbox.py)
batched             additive batched entry points ``assign_targets`` / ``detect``
sharding            image sharding across ranks + allgather of detections
synth               seeded synthetic GT / prediction generators
_lib                ctypes binding of the C-ABI in include/jabd_b200.h

``R/`` = the reference tree ``JABD2080ti/``.
"""
