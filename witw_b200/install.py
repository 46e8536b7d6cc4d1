"""Drop-in installation onto the reference's modules.

The reference has no plugin mechanism: train()/test() in model/cvig_fov.py and
tools/heatmap/heatmap.py look up ``correlation``, ``crop_overhead``, ``l2_distance``,
``PolarTransform`` and ``bilinear_interpolate`` as module attributes at call time
(cvig_fov.py:396, 450-453, 500, 537-538, 547-549; heatmap.py:106, 172-175).  Rebinding those
attributes on the imported module therefore swaps the hot path for every caller:

    import cvig_fov, witw_b200
    witw_b200.install(cvig_fov)        # correlation / crop_overhead / l2_distance / PolarTransform now run the B200 kernels
                                       # (the transform chain needs num_workers=0 or a spawn context: see install())
"""
from . import ops

_REBOUND = ("bilinear_interpolate", "correlation", "crop_overhead", "l2_distance")
# PolarTransform sits in the dataset's transform chain (cvig_fov.py:396, 500), which the reference runs inside forked
# DataLoader workers; the drop-in works there only with num_workers=0 or a spawn context (it raises a clear error
# otherwise), so it can be left out: install(module, polar=False)
_POLAR = ("PolarTransform",)
# train() picks its loss up as a module global too (cvig_fov.py:413); rebound where the module defines it
_REBOUND_IF_PRESENT = ("triplet_loss",)
_ADDED = ("match", "match_distance", "evaluate_ranks", "recall_from_ranks", "heatmap_scores", "polar_transform")


# the dataset transforms upstream of PolarTransform (cvig_fov.py:100-149); rebound only on request because the reference runs
# them inside forked DataLoader workers, where CUDA is not available (use num_workers=0 or a spawn context)
_TRANSFORMS = ("Resize", "ImageNormalization")


def install(module, transforms=False, polar=True):
    """Rebind the hot-path names of a reference module (cvig_fov / cvig_semantic) to witw_b200.

    polar=False leaves the reference's CPU ``PolarTransform`` in place (for loaders with forked workers; apply
    ``witw_b200.polar_transform`` to the batch after the loader instead).  transforms=True also rebinds ``Resize`` and
    ``ImageNormalization``, which sit in the same chain.
    Returns the dict of replaced originals (also kept on the module as ``_witw_b200_originals``).
    """
    originals = {}
    for name in _REBOUND + (_POLAR if polar else ()) + (_TRANSFORMS if transforms else ()):
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
    for name in _REBOUND + _POLAR + _REBOUND_IF_PRESENT + _TRANSFORMS + _ADDED:
        if name in originals:
            setattr(module, name, originals[name])
        elif name in _ADDED and hasattr(module, name):
            delattr(module, name)
    del module._witw_b200_originals
