"""oracle/refstub.py -- TEST INFRASTRUCTURE.  Imports the REAL reference Python (read-only checkout at
/root/reference) in this container by stubbing its absent third-party dependencies.

Used only by tests/golden/make_golden.py to generate golden fixtures (the checkout does not exist on the
GPU box).  Nothing here is imported by the product.  Stubs: pytorch_lightning, mmcv.Config (attr dict),
mmdet / mmdet3d builders, spconv (names only), pyquaternion, skimage, nuscenes, numba-free helpers and the
three prebuilt CUDA extension module names.  `Tensor.cuda` is patched to the identity on CPU-only hosts
because the reference calls `.cuda()` unconditionally (e.g. ..._distill_lidar.py:127).
"""
import os
import sys
import types

import torch
from torch import nn

REFERENCE_ROOT = os.environ.get("UNIDISTILL_REFERENCE_ROOT", "/root/reference")


class _AttrDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {}, **kw)
        for k, v in d.items():
            self[k] = self._wrap(v)

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, _AttrDict):
            return cls(v)
        if isinstance(v, list):
            return [cls._wrap(x) for x in v]
        return v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = self._wrap(v)


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Dummy(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()

    def init_weights(self):
        return None

    def forward(self, x):
        return x


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "unidistill"))


class _Any(object):
    """Registry / function stand-in: attribute access and calls succeed; usable as a decorator."""

    def __call__(self, *a, **k):
        if len(a) == 1 and not k and (isinstance(a[0], type) or callable(a[0])):
            return a[0]
        return _Any()

    def __getattr__(self, name):
        return _Any()


class _StubModule(types.ModuleType):
    """Module whose missing attributes materialise on demand: CamelCase -> empty class, else _Any()."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        if name[0].isupper() and not name.isupper():
            val = type(name, (object,), {})
        else:
            val = _Any()
        setattr(self, name, val)
        return val


_STUB_ROOTS = ("mmdet", "mmdet3d", "mmcv", "mmseg", "nuscenes", "skimage", "pyquaternion", "pytorch_lightning",
               "spconv", "easydict", "open3d")


class _StubFinder(object):
    """meta_path finder: any import below the stubbed third-party roots yields a _StubModule."""

    def find_spec(self, fullname, path=None, target=None):
        import importlib.machinery

        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        return None


AttrDict = _AttrDict


def install(overlay=False):
    """Puts the reference on sys.path (first) and registers the third-party stubs.  Must run in a process
    that has NOT imported this repo's own `unidistill` mirror.
    overlay=True: the repo's `unidistill` mirror and `spconv` shim stay in front (they must already be on sys.path) and
    the reference is reached through UNIDISTILL_REFERENCE_ROOT (unidistill/_overlay.py) -- the deployment of
    INTEGRATION.md section 1; only the packages absent from this image are stubbed."""
    if not available():
        raise RuntimeError("reference checkout not found at %s" % REFERENCE_ROOT)
    if overlay:
        os.environ["UNIDISTILL_REFERENCE_ROOT"] = REFERENCE_ROOT
    else:
        if "unidistill" in sys.modules and not sys.modules["unidistill"].__file__.startswith(REFERENCE_ROOT):
            raise RuntimeError("the mirror package is already imported in this process")
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.append(_StubFinder())

    import importlib

    pl = importlib.import_module("pytorch_lightning")
    pl.LightningModule = nn.Module
    pl.seed_everything = lambda s, **k: torch.manual_seed(s)
    importlib.import_module("pytorch_lightning.core").LightningModule = nn.Module
    importlib.import_module("mmcv").Config = _AttrDict
    importlib.import_module("mmdet.models").build_backbone = lambda cfg: _Dummy()
    importlib.import_module("mmdet3d.models").build_neck = lambda cfg: _Dummy()

    class _SparseModule(nn.Module):
        pass

    if overlay:
        _mod("unidistill.utils.det3d_utils.roiaware_pool3d_cuda")      # data augmentation only; no mirror module
    else:
        spt = importlib.import_module("spconv.pytorch")
        spt.SparseModule = _SparseModule
        spt.SparseSequential = nn.Sequential
        importlib.import_module("spconv.core").ConvAlgo = types.SimpleNamespace(Native=0)
        for so in ("unidistill.layers.blocks_3d.mmdet3d.voxel_pooling_ext",
                   "unidistill.layers.head.det3d.generate_proposals.iou3d_nms_cuda",
                   "unidistill.utils.det3d_utils.roiaware_pool3d_cuda"):
            _mod(so)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    return _AttrDict


def load_distill_module(script="BEVFusion_nuscenes_centerhead_camera_exp_distill_lidar"):
    """Imports one of the reference's distill experiment scripts (module-level loss functions)."""
    import importlib

    return importlib.import_module(
        "unidistill.exps.multisensor_fusion.nuscenes.BEVFusion." + script)
