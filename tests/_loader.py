"""Loads the product binding (directory name has hyphens) as module `dogm_b200`, and the oracle wrapper as `dogm_oracle`.

Only tests/, bench.py and __graft_entry__.py import the oracle; the product never does."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "dynamic-occupancy-grid-map_b200")


def load_dogm_b200():
    if "dogm_b200" in sys.modules:
        return sys.modules["dogm_b200"]
    spec = importlib.util.spec_from_file_location(
        "dogm_b200", os.path.join(PKG_DIR, "__init__.py"), submodule_search_locations=[PKG_DIR]
    )
    mod = importlib.util.module_from_spec(spec)
    sys.modules["dogm_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


def load_oracle():
    if "dogm_oracle" in sys.modules:
        return sys.modules["dogm_oracle"]
    spec = importlib.util.spec_from_file_location("dogm_oracle", os.path.join(ROOT, "oracle", "oracle.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["dogm_oracle"] = mod
    spec.loader.exec_module(mod)
    return mod


def load_ref():
    """Wrapper of oracle/_ref (the reference's own CUDA code); check `.available()` before use."""
    if "dogm_ref" in sys.modules:
        return sys.modules["dogm_ref"]
    spec = importlib.util.spec_from_file_location("dogm_ref", os.path.join(ROOT, "oracle", "ref.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["dogm_ref"] = mod
    spec.loader.exec_module(mod)
    return mod
