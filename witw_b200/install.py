"""Drop-in installation onto the reference's modules.

The reference has no plugin mechanism: train()/test() in model/cvig_fov.py and
tools/heatmap/heatmap.py look up ``correlation``, ``crop_overhead``, ``l2_distance``,
``PolarTransform`` and ``bilinear_interpolate`` as module attributes at call time
(cvig_fov.py:396, 450-453, 500, 537-538, 547-549; heatmap.py:106, 172-175).  Rebinding those
attributes on the imported module therefore swaps the hot path for every caller:

    import cvig_fov, witw_b200
    witw_b200.install(cvig_fov)        # cvig_fov.test() now runs the B200 kernels
"""
from . import ops

_REBOUND = ("bilinear_interpolate", "PolarTransform", "correlation", "crop_overhead", "l2_distance")
# train() picks its loss up as a module global too (cvig_fov.py:413); rebound where the module defines it
_REBOUND_IF_PRESENT = ("triplet_loss",)
_ADDED = ("match", "match_distance", "evaluate_ranks", "recall_from_ranks", "heatmap_scores", "polar_transform")


# the dataset transforms upstream of PolarTransform (cvig_fov.py:100-149); rebound only on request because the reference runs
# them inside forked DataLoader workers, where CUDA is not available (use num_workers=0 or a spawn context)
_TRANSFORMS = ("Resize", "ImageNormalization")


def install(module, transforms=False):
    """Rebind the hot-path names of a reference module (cvig_fov / cvig_semantic) to witw_b200.

    transforms=True also rebinds ``Resize`` and ``ImageNormalization``.
    Returns the dict of replaced originals (also kept on the module as ``_witw_b200_originals``).
    """
    originals = {}
    for name in _REBOUND + (_TRANSFORMS if transforms else ()):
        if not hasattr(module, name):
            raise AttributeError("install: %s has no attribute %r -- not a WITW cvig module?" % (getattr(module, "__name__", module), name))
        originals[name] = getattr(module, name)
        setattr(module, name, getattr(ops, name))
    for name in _REBOUND_IF_PRESENT:
        if hasattr(module, name):
            originals[name] = getattr(module, name)
            setattr(module, name, getattr(ops, name))
    for name in _ADDED:
        if hasattr(module, name):
            originals[name] = getattr(module, name)
        setattr(module, name, getattr(ops, name))
    module._witw_b200_originals = originals
    return originals


def uninstall(module):
    """Undo install()."""
    originals = getattr(module, "_witw_b200_originals", None)
    if originals is None:
        return
    for name in _REBOUND + _REBOUND_IF_PRESENT + _TRANSFORMS + _ADDED:
        if name in originals:
            setattr(module, name, originals[name])
        elif name in _ADDED and hasattr(module, name):
            delattr(module, name)
    del module._witw_b200_originals
