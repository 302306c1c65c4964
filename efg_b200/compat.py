"""Import aliases so code written against the reference's module names resolves to this package.

``install()`` registers (only where no real module of that name is importable):
    spconv, spconv.pytorch            -> efg_b200.spconv                  (sparse_net.py:6-11)
    efg._C                            -> efg_b200._C                      (operators/*.py import it)
    efg.operators, efg.modeling.operators -> efg_b200.operators           (VD/modules/box_attention.py:7
                                                                           imports the latter, which does
                                                                           not exist in the reference tree)
    torch._six                        -> shim with string_classes         (VD/modules/utils.py:11; removed in torch 2)
"""
import importlib
import sys
import types


def _missing(name):
    if name in sys.modules:
        return False
    try:
        return importlib.util.find_spec(name) is None
    except (ImportError, ValueError, AttributeError):
        return True


def _ensure_package(name):
    if name in sys.modules:
        return sys.modules[name]
    mod = types.ModuleType(name)
    mod.__path__ = []  # mark as package
    sys.modules[name] = mod
    if "." in name:
        parent, child = name.rsplit(".", 1)
        setattr(_ensure_package(parent), child, mod)
    return mod


def install(force=False):
    """Register the aliases.  A REAL `efg` package on sys.path (the reference checkout) is never shadowed: it is
    imported and only the pieces it lacks are attached to it — `efg._C` (the reference builds it as a CUDAExtension,
    setup.py:63-71; here the C-ABI library stands in) and `efg.modeling.operators` (imported by
    VD/modules/box_attention.py:7 but absent from the tree).  Without a real `efg`, a namespace stub carries the aliases."""
    from . import _C, operators, spconv
    from .spconv import pytorch as spconv_pytorch

    installed = []
    if force or _missing("spconv"):
        sys.modules["spconv"] = spconv
        sys.modules["spconv.pytorch"] = spconv_pytorch
        installed += ["spconv", "spconv.pytorch"]
    real_efg = None
    if not force and not _missing("efg") and "efg" not in sys.modules:
        real_efg = importlib.import_module("efg")
    elif "efg" in sys.modules and getattr(sys.modules["efg"], "__file__", None):
        real_efg = sys.modules["efg"]
    if real_efg is not None:
        if _missing("efg._C"):
            sys.modules["efg._C"] = _C
            real_efg._C = _C
            installed.append("efg._C")
        try:
            modeling = importlib.import_module("efg.modeling")
        except ImportError:
            modeling = _ensure_package("efg.modeling")
        if _missing("efg.modeling.operators"):
            sys.modules["efg.modeling.operators"] = operators
            modeling.operators = operators
            installed.append("efg.modeling.operators")
    else:
        efg = _ensure_package("efg")
        _ensure_package("efg.modeling")
        for alias, target in (("efg._C", _C), ("efg.operators", operators), ("efg.modeling.operators", operators)):
            sys.modules[alias] = target
            parent, child = alias.rsplit(".", 1)
            setattr(sys.modules[parent], child, target)
            installed.append(alias)
        efg.__dict__.setdefault("_C", _C)
    if _missing("torch._six"):
        six = types.ModuleType("torch._six")
        six.string_classes = (str, bytes)
        sys.modules["torch._six"] = six
        installed.append("torch._six")
    return installed
