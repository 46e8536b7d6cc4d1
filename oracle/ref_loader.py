"""TEST INFRASTRUCTURE ONLY.

Imports the *unmodified* reference modules from /root/reference/model so the
oracle restatement (oracle/witw_oracle.py) can be pinned against them and so
tests/golden/make_golden.py can freeze golden vectors.  The reference tree does
not exist on the GPU box: nothing in the gpu tests / smoke / bench calls this.

Recipe follows SURVEY.md section 8(c): stub the two imports that are missing in
this image (skimage, tifffile; only the dataset classes touch them) and patch
torch.hub.load (no network) to build an un-pretrained torchvision VGG16.
"""
import os
import sys
import types

REFERENCE_MODEL_DIR = os.environ.get("WITW_REFERENCE_DIR", "/root/reference/model")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_MODEL_DIR, "cvig_fov.py"))


def load(name: str = "cvig_fov"):
    """Return the reference module ``name`` (cvig_fov | cvig_semantic | cvig_baseline), forced to CPU."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_MODEL_DIR)
    import torch
    import torchvision

    for m in ("skimage", "skimage.io", "tifffile"):
        if m not in sys.modules:
            sys.modules[m] = types.ModuleType(m)
    sys.modules["skimage"].io = sys.modules["skimage.io"]
    if REFERENCE_MODEL_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_MODEL_DIR)

    def _hub_load(repo, model, pretrained=True, **kw):
        return getattr(torchvision.models, model)(weights=None)

    torch.hub.load = _hub_load
    mod = __import__(name)
    if hasattr(mod, "device"):
        mod.device = torch.device("cpu")
    return mod
