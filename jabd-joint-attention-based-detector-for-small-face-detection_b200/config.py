"""Prior-box configuration contract.

The hot path reads exactly four keys of the reference's cfg dicts
(R/utils/anchors.py:11-13, R/predict.py:167): ``min_sizes``, ``steps``,
``variance`` and ``clip``.  The tables below restate those four keys for the
eight configurations the reference ships (R/utils/config.py:1-152); any dict
with the same keys -- including the reference's own -- is accepted by every
entry point in this package.
"""

def _cfg(name, min_sizes, steps, clip=False, variance=(0.1, 0.2)):
    return {
        "name": name,
        "min_sizes": [list(m) for m in min_sizes],
        "steps": list(steps),
        "variance": list(variance),
        "clip": clip,
    }

_P3 = ((16, 32), (64, 128), (256, 512))
_P4 = ((8, 16), (32, 64), (64, 128), (256, 512))

cfg_mnet = _cfg("mobilenet0.25", _P3, (8, 16, 32))                       # R/utils/config.py:1-19
cfg_mnet_4 = _cfg("mobilenetV3", ((4, 12),) + _P3, (8, 16, 16, 32))       # :20-40 (repeated step)
cfg_re50 = _cfg("Resnet50", _P3, (8, 16, 32))                            # :42-55
cfg_re50_self = _cfg("Resnet50_self", _P4, (8, 16, 32, 64))               # :56-81
cfg_re152_ = _cfg("Resnet152", _P3, (8, 16, 32))                         # :82-93
cfg_re152 = _cfg("Resnet152", _P4, (4, 8, 16, 32))                        # :95-112
cfg_re101 = _cfg("Resnet101", ((32, 64), (64, 128), (256, 512), (240, 480)),
                 (8, 16, 32, 60))                                         # :113-131 (odd step 60)
cfg_re152_new = _cfg("Resnet152", _P4, (4, 8, 16, 32))                    # :132-152

ALL_CFGS = {
    "cfg_mnet": cfg_mnet, "cfg_mnet_4": cfg_mnet_4, "cfg_re50": cfg_re50,
    "cfg_re50_self": cfg_re50_self, "cfg_re152_": cfg_re152_, "cfg_re152": cfg_re152,
    "cfg_re101": cfg_re101, "cfg_re152_new": cfg_re152_new,
}


def num_priors(cfg, image_size):
    """Prior count for (H, W): sum over levels of ceil(H/s)*ceil(W/s)*len(min_sizes)."""
    h, w = int(image_size[0]), int(image_size[1])
    return sum((-(-h // s)) * (-(-w // s)) * len(m) for s, m in zip(cfg["steps"], cfg["min_sizes"]))
