"""Stand-ins for the reference's third-party dependencies that are absent from this image (Lightning, torchmetrics, the
geo stack ...), so that the reference's own pipeline modules can be IMPORTED and its `load_pipeline` can construct a
pipeline class around the B200 plug-ins.  Test infrastructure for tests/test_plugin_boundary.py only.

Any module below one of the STUBBED top-level packages that cannot be found is fabricated on import; attribute access on a
fabricated module returns a permissive dummy class (callable, subclassable, chainable).  `pytorch_lightning.LightningModule`
is a real `torch.nn.Module` subclass with the few hooks the reference's `Pipeline` touches."""
import importlib.abc
import importlib.machinery
import sys
import types
import typing

import torch

STUBBED = ("pytorch_lightning", "lightning", "gpustat", "torchmetrics", "cv2", "kornia", "rasterio", "rpcm", "utm",
           "matplotlib", "numba", "plyflatten", "pyntcloud", "fire", "osgeo", "srtm4", "affine", "torchvision", "PIL",
           "pyproj", "shapely", "skimage", "imageio", "open3d", "tensorboard", "geojson", "pandas_stub", "bs4", "requests_stub",
           "iio", "ransac", "s2p", "tifffile", "pyquaternion", "trimesh", "seaborn", "glob2", "datadings", "natsort", "pymap3d",
           "pycocotools")


class _Anything:
    """callable, subclassable, attribute-chainable placeholder"""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (_Anything,)


class LightningModule(torch.nn.Module):
    """the slice of pl.LightningModule the reference's Pipeline uses at construction time"""
    current_epoch = 0
    global_step = 0

    def log(self, *a, **k):
        pass

    def log_dict(self, *a, **k):
        pass

    def save_hyperparameters(self, *a, **k):
        pass


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        if name in ("LightningModule",):
            return LightningModule
        val = type(name, (_Anything,), {})
        setattr(self, name, val)
        return val


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in STUBBED:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_installed = []


def install():
    """append the stub finder (real modules keep priority) and patch the one torch symbol the reference imports that newer
    torch versions dropped (torch.utils.data.dataset.T_co, framework/datasets.py:6)"""
    if _installed:
        return
    f = _Finder()
    sys.meta_path.append(f)
    _installed.append(f)
    import torch.utils.data.dataset as tds
    if not hasattr(tds, "T_co"):
        tds.T_co = typing.TypeVar("T_co", covariant=True)
        _installed.append("T_co")


def uninstall():
    import torch.utils.data.dataset as tds
    for x in _installed:
        if x == "T_co":
            del tds.T_co
        else:
            sys.meta_path.remove(x)
    _installed.clear()
    for name in [n for n in sys.modules if n.split(".")[0] in STUBBED and isinstance(sys.modules[n], _StubModule)]:
        del sys.modules[name]
