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


def _main(argv):
    """``python -m rca_b200_loader <submodule> [args...]`` runs ``realtime_codec_agent_b200.<submodule>.main(args)`` —
    the ``python -m package.module`` form for a package whose directory name is not an identifier, e.g.
    ``torchrun --nproc-per-node 8 -m rca_b200_loader audio_to_codes --audio_path raw --codes_path codes``."""
    import importlib
    if not argv or argv[0] in ("-h", "--help"):
        print(_main.__doc__)
        return 0 if argv else 2
    mod = importlib.import_module(f"{PKG_NAME}.{argv[0]}")
    if not hasattr(mod, "main"):
        print(f"{PKG_NAME}.{argv[0]} has no main()", file=sys.stderr)
        return 2
    mod.main(argv[1:])
    return 0


if __name__ == "__main__":
    sys.exit(_main(sys.argv[1:]))
