"""Import shim: the product package lives in the directory
``audiovidsum-a-multi-modal-approach-to-video-summarization_b200/`` (a name Python
cannot import directly because of the hyphens).  ``import avsum_b200`` loads that
directory as the package ``avsum_b200`` so that ``avsum_b200.models.av_model`` etc.
resolve normally."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                        "audiovidsum-a-multi-modal-approach-to-video-summarization_b200")
_spec = importlib.util.spec_from_file_location(
    "avsum_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["avsum_b200"] = _mod
_spec.loader.exec_module(_mod)
