"""Short import alias for the package directory
``decentralized-multi-target-cross-domain-recommendation-for-multi-organization-collaborations_b200/``.

The directory name is not a valid Python identifier, so it is loaded here under
the canonical module name ``dmtcdr_b200`` (relative imports inside the package
resolve under that name; nothing imports it under the long name).
"""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(
    os.path.dirname(os.path.abspath(__file__)),
    "decentralized-multi-target-cross-domain-recommendation-for-multi-organization-collaborations_b200",
)


def _load():
    spec = importlib.util.spec_from_file_location(
        "dmtcdr_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR]
    )
    module = importlib.util.module_from_spec(spec)
    sys.modules["dmtcdr_b200"] = module
    spec.loader.exec_module(module)
    return module


_load()
