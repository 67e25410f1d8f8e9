"""Import shim for the UNMODIFIED reference at /root/reference (test infrastructure only).

TEST INFRASTRUCTURE - NOT PRODUCT CODE.  Only tests/, tests/golden/make_golden.py and
oracle/ may import this module.  It exists only in the authoring container: /root/reference
does not travel to the GPU box, so nothing under `-m gpu`, smoke() or bench.py uses it.

The reference (p-mc-grath/DMMFODS) imports symbols that modern torchvision removed
(`torchvision.models.densenet.model_urls`, `torchvision.models.utils`;
Dense_U_Net_lidar.py:9-10) and packages that are not installed (easydict, tensorflow,
waymo_open_dataset; Dense_U_Net_lidar_helper.py:4,9,15-18).  We stub those in sys.modules
and import the reference's own source files from where they lie (recipe: SURVEY.md App. C).
"""
import importlib.machinery as _im
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DMMFODS_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(
        REFERENCE_ROOT, "dmmfods", "graphs", "models", "Dense_U_Net_lidar.py"))


class _EasyDict(dict):
    """Stand-in for easydict.EasyDict (helper:9,223): attribute AND item access, recursive."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            setattr(self, k, v)

    def __setattr__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, _EasyDict):
            v = _EasyDict(v)
        elif isinstance(v, (list, tuple)):
            v = type(v)(_EasyDict(x) if isinstance(x, dict) else x for x in v)
        super().__setattr__(k, v)
        super().__setitem__(k, v)

    __setitem__ = __setattr__


_installed = False


def install():
    """Install the stubs (idempotent).  ORDER MATTERS: real torch/torchvision/tensorboard first."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    import torch  # noqa: F401
    import torchvision  # noqa: F401
    try:
        import torch.utils.tensorboard  # noqa: F401  (must precede the tensorflow stub)
    except Exception:
        pass
    names = ["tensorflow", "waymo_open_dataset", "waymo_open_dataset.utils",
             "waymo_open_dataset.utils.range_image_utils",
             "waymo_open_dataset.utils.transform_utils",
             "waymo_open_dataset.utils.frame_utils", "waymo_open_dataset.dataset_pb2"]
    for n in names:
        if n in sys.modules:
            continue
        m = types.ModuleType(n)
        m.__spec__ = _im.ModuleSpec(n, None)
        sys.modules[n] = m
    for n in names:
        if "." in n:
            parent, child = n.rsplit(".", 1)
            setattr(sys.modules[parent], child, sys.modules[n])
    if "easydict" not in sys.modules:
        ed = types.ModuleType("easydict")
        ed.__spec__ = _im.ModuleSpec("easydict", None)
        ed.EasyDict = _EasyDict
        sys.modules["easydict"] = ed
    import torchvision.models.densenet as d
    if not hasattr(d, "model_urls"):
        d.model_urls = {}
    if "torchvision.models.utils" not in sys.modules:
        import torch
        u = types.ModuleType("torchvision.models.utils")
        u.__spec__ = _im.ModuleSpec("torchvision.models.utils", None)
        u.load_state_dict_from_url = torch.hub.load_state_dict_from_url
        sys.modules["torchvision.models.utils"] = u
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def ref_modules():
    """Returns (model_module, helper_module) of the unmodified reference."""
    install()
    import dmmfods.graphs.models.Dense_U_Net_lidar as model_mod
    import dmmfods.utils.Dense_U_Net_lidar_helper as helper_mod
    return model_mod, helper_mod


def ref_config(stream_2_in_channels=1, concat_before_block_num=2, **model_overrides):
    _, helper = ref_modules()
    cfg = helper.get_config("/nonexistent")
    cfg.model.stream_2_in_channels = stream_2_in_channels
    cfg.model.concat_before_block_num = concat_before_block_num
    for k, v in model_overrides.items():
        setattr(cfg.model, k, v)
    return cfg
