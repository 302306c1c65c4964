"""Run the REFERENCE's own Python (the `efg` package and the 3-D playground model files, unmodified, from
/root/reference) inside this container, so that goldens for the model-level code can be produced by the reference
itself and so that the reference's model files can be executed over this repo's `spconv` / `efg._C` /
`efg.operators` surface.

Test infrastructure only; needs /root/reference (absent on the GPU box: everything that uses this module skips there).

What is provided around the unmodified reference sources:
  * the real `efg` package is imported from /root/reference (NOT shadowed by a stub package);
  * third-party modules that are absent from this image and that the model files never execute
    (portalocker, pycocotools, termcolor, pyquaternion, ...) become permissive stub modules;
  * `omegaconf` becomes a 40-line stand-in (`OmegaConf.create / to_container`, attribute dicts) — the model
    files only read config values;
  * `torch._six.string_classes` (removed in torch 2, imported by VD/modules/utils.py:11);
  * `spconv` / `spconv.pytorch`  -> the module passed to install() (oracle.spconv_cpu on the CPU,
    efg_b200.spconv on a GPU);
  * `efg._C`                      -> the module passed to install() (efg_b200._C; a stub on the CPU);
  * `efg.modeling.operators`      -> a module with `BoxAttnFunction` (VD/modules/box_attention.py:7 imports this
    path, which does not exist in the reference tree: upstream it is an alias of efg.operators).  On the CPU the
    function is the reference's own torch twin `ms_deform_attn_core_pytorch` (efg/operators/ms_deform_attn.py:55-76).
"""
import contextlib
import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types

REF = "/root/reference"
PLAY = os.path.join(REF, "playground/detection.3d/waymo")
VD_DIR = os.path.join(PLAY, "conquer/VoxelDETR.waymo.res18.p3.box_only_with_3cat.bs6.epoch6")
CQ_DIR = os.path.join(PLAY, "conquer/ConQueR.waymo.res18.p3.dn3.tau07.noised_only.bs6.epoch6")
CP_DIR = os.path.join(PLAY, "center_point/centerpoint.waymo.voxelnet.gt_aug.ds_sample.onecycle.adam.bs48.36e")

# top-level third-party packages that may be stubbed when they are not importable
STUBBABLE = {"portalocker", "pycocotools", "termcolor", "pyquaternion", "nuscenes", "waymo_open_dataset", "tomark",
             "easydict", "cv2", "shapely", "tensorboard", "tensorflow", "colorama", "lap", "motmetrics", "fire",
             "seaborn", "matplotlib", "skimage", "open3d", "mayavi", "panopticapi", "lvis", "cityscapesscripts",
             "fvcore", "iopath", "timm", "numba_stub_never"}


# stubbed even though a module of that name is installed (the image's cv2 raises on import: cv2.dnn.DictValue)
FORCE_STUB = {"cv2"}


class _ForceStubFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in FORCE_STUB:
            return importlib.machinery.ModuleSpec(fullname, _StubFinder(), is_package=True)
        return None


def available():
    return os.path.isdir(os.path.join(REF, "efg"))


class _Anything:
    """Attribute / call sink for stubbed third-party names that are referenced at import time only."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:  # used as a decorator
            return a[0]
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        val = type(name, (object,), {"__init__": lambda self, *a, **k: None}) if name[:1].isupper() else _Anything()
        setattr(self, name, val)
        return val


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in STUBBABLE:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        mod = _StubModule(spec.name)
        mod.__path__ = []
        return mod

    def exec_module(self, module):
        pass


class AttrDict(dict):
    """Config node: dict with attribute access (what the model files need from OmegaConf's DictConfig)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def __deepcopy__(self, memo):
        import copy
        return AttrDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


def to_cfg(obj):
    if isinstance(obj, dict):
        return AttrDict({k: to_cfg(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)):
        return [to_cfg(v) for v in obj]
    return obj


def _to_container(obj, resolve=False):
    if isinstance(obj, dict):
        return {k: _to_container(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [_to_container(v) for v in obj]
    return obj


def _omegaconf_module():
    m = types.ModuleType("omegaconf")

    class OmegaConf:
        create = staticmethod(lambda obj=None: to_cfg(obj if obj is not None else {}))
        to_container = staticmethod(_to_container)
        to_yaml = staticmethod(lambda cfg, **k: repr(cfg))
        is_config = staticmethod(lambda obj: isinstance(obj, AttrDict))
        merge = staticmethod(lambda *cfgs: to_cfg({k: v for c in cfgs for k, v in c.items()}))

    m.OmegaConf = OmegaConf
    m.DictConfig = AttrDict
    m.ListConfig = list
    return m


_installed = {}


def install(spconv_module=None, c_module=None, box_attn_function=None):
    """Make `import efg...` / `from voxel_detr import VoxelDETR` work on the reference sources.  Returns nothing;
    call `uninstall()` (or use `reference_modules(...)`) to restore sys.modules / sys.path."""
    assert available(), "/root/reference is not present"
    assert not _installed, "ref_env.install() is already active"
    _installed["path"] = list(sys.path)
    _installed["modules"] = dict(sys.modules)
    _installed["finder"] = _StubFinder()
    # a repo-side alias package named `efg` (efg_b200.compat) must not shadow the real one
    for name in [n for n in sys.modules if n == "efg" or n.startswith("efg.") or n == "spconv" or n.startswith("spconv.")]:
        del sys.modules[name]
    sys.path.insert(0, REF)
    sys.meta_path.append(_installed["finder"])
    _installed["force"] = _ForceStubFinder()
    sys.meta_path.insert(0, _installed["force"])
    import collections
    import collections.abc
    for name in ("Mapping", "MutableMapping", "Sequence", "Iterable", "Callable"):  # removed from `collections` in py3.10
        if not hasattr(collections, name):                                         # (efg/engine/hooks.py:4)
            setattr(collections, name, getattr(collections.abc, name))
    try:
        importlib.import_module("omegaconf")
    except ImportError:
        sys.modules["omegaconf"] = _omegaconf_module()
    try:
        importlib.import_module("torch._six")
    except ImportError:
        six = types.ModuleType("torch._six")
        six.string_classes = (str, bytes)
        sys.modules["torch._six"] = six
    import efg  # the REAL package (/root/reference/efg/__init__.py)

    assert os.path.realpath(efg.__file__).startswith(REF), efg.__file__
    if c_module is None:
        c_module = _StubModule("efg._C")
    sys.modules["efg._C"] = c_module
    efg._C = c_module
    if spconv_module is not None:
        sys.modules["spconv"] = spconv_module
        sys.modules["spconv.pytorch"] = getattr(spconv_module, "pytorch", spconv_module)
    if box_attn_function is None:
        box_attn_function = reference_box_attn_function()
    ops_alias = types.ModuleType("efg.modeling.operators")
    ops_alias.BoxAttnFunction = box_attn_function
    sys.modules["efg.modeling.operators"] = ops_alias
    import efg.modeling as efg_modeling  # namespace package inside the reference

    efg_modeling.operators = ops_alias
    # CP/voxelnet.py:9 imports `efg.data.augmentations3d._dict_select`, a module the reference tree no longer has;
    # the function lives in efg/data/utils/misc.py:1-9 (SURVEY.md §8b import closure)
    import efg.data.utils.misc as _misc

    aug3d = types.ModuleType("efg.data.augmentations3d")
    aug3d._dict_select = _misc._dict_select
    sys.modules["efg.data.augmentations3d"] = aug3d


def uninstall():
    if not _installed:
        return
    sys.meta_path.remove(_installed["finder"])
    sys.meta_path.remove(_installed["force"])
    sys.path[:] = _installed["path"]
    # drop what install() put in place (the reference's efg package, the aliases and the stubs); real third-party
    # modules imported meanwhile (cv2, numba, ...) stay loaded: C extensions do not survive a second import
    for name in list(sys.modules):
        top = name.split(".")[0]
        if name in _installed["modules"]:
            continue
        mod = sys.modules[name]
        if top in ("efg", "spconv") or isinstance(mod, _StubModule) or top in STUBBABLE or \
                (top == "omegaconf" and not hasattr(mod, "__file__")) or name == "torch._six":
            del sys.modules[name]
    for name, mod in _installed["modules"].items():
        if name.split(".")[0] in ("efg", "spconv"):
            sys.modules[name] = mod
    _installed.clear()


def reference_box_attn_function():
    """`BoxAttnFunction.apply(value, shapes, level_start, loc, attn, im2col_step)` implemented by the reference's
    own torch twin of the CUDA kernel (efg/operators/ms_deform_attn.py:55-76); autograd differentiates it."""
    spec = importlib.util.spec_from_file_location("_ref_ms_deform_attn", os.path.join(REF, "efg/operators/ms_deform_attn.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)

    class BoxAttnFunction:
        @staticmethod
        def apply(value, spatial_shapes, level_start_index, sampling_locations, attention_weights, im2col_step=64):
            b, lq, h, l = attention_weights.shape[:4]
            attn = attention_weights.reshape(b, lq, h, l, -1)
            return mod.ms_deform_attn_core_pytorch(value, spatial_shapes, sampling_locations, attn)

    return BoxAttnFunction


@contextlib.contextmanager
def cuda_calls_stay_on_cpu():
    """The ConQueR playground hard-codes `.cuda()` / `.to("cuda")` on freshly created tensors (CQ/cdn.py:20-94,
    CQ/losses.py:161-167).  To run those unmodified files on the CPU of this container the two calls become no-ops
    for the duration of the block (an environment shim, like the missing third-party modules above)."""
    import torch

    orig_cuda, orig_to = torch.Tensor.cuda, torch.Tensor.to

    def to(self, *a, **k):
        a = tuple(x for x in a if not (isinstance(x, str) and x.startswith("cuda")) and
                  not (isinstance(x, torch.device) and x.type == "cuda"))
        if isinstance(k.get("device"), (str, torch.device)) and str(k["device"]).startswith("cuda"):
            k = {kk: v for kk, v in k.items() if kk != "device"}
        return orig_to(self, *a, **k) if (a or k) else self

    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.Tensor.to = to
    try:
        yield
    finally:
        torch.Tensor.cuda, torch.Tensor.to = orig_cuda, orig_to


@contextlib.contextmanager
def playground(exp_dir, **kw):
    """`with playground(VD_DIR): from voxel_detr import VoxelDETR` — cwd-style imports of one experiment directory
    (the reference's CLI chdir's into it, cli/main.py:120,144).  Experiment-local module names (heads, losses,
    transformer, modules.*, ...) are purged on exit so a second experiment can be loaded afterwards."""
    install(**kw)
    sys.path.insert(0, exp_dir)
    before = set(sys.modules)
    try:
        yield
    finally:
        for name in set(sys.modules) - before:
            f = getattr(sys.modules[name], "__file__", None) or ""
            if f.startswith(exp_dir):
                del sys.modules[name]
        uninstall()
