"""Import the *real* reference package from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY.  ``/root/reference`` does not exist on the GPU box, so this
module is used solely by ``oracle/make_golden.py`` (fixture generation) and by the
``not gpu`` test that re-validates the restatement against the live reference when the
tree happens to be mounted.

The reference's ``ultralytics.nn`` package imports seven third-party packages that are
not installed in this image; none of them is touched by the
post-processing path, so they are replaced by inert ``MagicMock`` modules.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import os
import sys
from unittest.mock import MagicMock

REFERENCE_ROOT = "/root/reference"
_STUBBED = ("timm", "pywt", "fairscale", "thop", "basicsr", "fvcore", "antialiased_cnns")


class _InertFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path=None, target=None):
        if name.split(".")[0] in _STUBBED:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        mod = MagicMock(name=spec.name)
        mod.__path__ = []
        mod.__spec__ = spec
        mod.__name__ = spec.name
        return mod

    def exec_module(self, module):
        return None


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ultralytics"))


_loaded = None


def load_reference():
    """Return a namespace with the reference's hot-path symbols (imports once)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"{REFERENCE_ROOT} is not mounted; the live reference only exists in the build container")
    os.environ.setdefault("YOLO_CONFIG_DIR", "/tmp/ypb_ref_cfg")
    os.makedirs(os.environ["YOLO_CONFIG_DIR"], exist_ok=True)
    sys.dont_write_bytecode = True
    if not any(isinstance(f, _InertFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _InertFinder())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from ultralytics.nn.modules import head as ref_head
        from ultralytics.utils import metrics as ref_metrics
        from ultralytics.utils import nms as ref_nms
        from ultralytics.utils import ops as ref_ops
        from ultralytics.utils import tal as ref_tal

    class _NS:
        head = ref_head
        nms = ref_nms
        tal = ref_tal
        metrics = ref_metrics
        ops = ref_ops
        Detect = ref_head.Detect
        OBB = ref_head.OBB
        non_max_suppression = staticmethod(ref_nms.non_max_suppression)
        TorchNMS = ref_nms.TorchNMS

    _loaded = _NS
    return _loaded
