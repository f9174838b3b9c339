"""Registers the hyphen-named package directory ``realtime-codec-agent_b200/`` as the importable
module ``realtime_codec_agent_b200`` (no symlink, no install step).  ``import rca_b200_loader``
once, then ``import realtime_codec_agent_b200`` works anywhere in the process."""
import importlib.util
import os
import sys

PKG_NAME = "realtime_codec_agent_b200"
PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "realtime-codec-agent_b200")


def load():
    if PKG_NAME in sys.modules:
        return sys.modules[PKG_NAME]
    spec = importlib.util.spec_from_file_location(
        PKG_NAME, os.path.join(PKG_DIR, "__init__.py"), submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[PKG_NAME] = mod
    spec.loader.exec_module(mod)
    return mod


package = load()
