"""Import alias and launcher of the package in `soft-intro-vae-pytorch_b200/` (a directory name Python cannot import by name).

    import sivae_b200                      -> the package (sivae_b200.lib, .engine, .train_soft_intro_vae, ...)
    python sivae_b200.py SCRIPT [ARGS...]  -> run an UNMODIFIED reference script (soft_intro_vae/main.py,
                                              soft_intro_vae_bootstrap/main.py, soft_intro_vae_2d/main.py) with the drop-in
                                              modules of this package resolving `from train_soft_intro_vae import ...`

Why a launcher: `python main.py` puts the SCRIPT's directory at sys.path[0], ahead of PYTHONPATH, so the reference's own
train_soft_intro_vae.py next to main.py would win the import.  The launcher runs the script with runpy and the search order
[this package's directory, the script's directory (for the reference's dataset.py / metrics/), the rest].  Equivalent without
the launcher:  PYTHONSAFEPATH=1 PYTHONPATH=<repo>/soft-intro-vae-pytorch_b200:<reference>/soft_intro_vae python main.py ...
(python >= 3.11; `python -P` is the same switch).
"""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
_PKG = "soft-intro-vae-pytorch_b200"


def _package():
    if _ROOT not in sys.path:
        sys.path.insert(0, _ROOT)
    return importlib.import_module(_PKG)


def launch(script, argv):
    """run `script` as __main__ with the drop-in modules first on the module path"""
    import runpy
    script = os.path.abspath(script)
    pkg_dir = os.path.join(_ROOT, _PKG)
    # drop the launcher's own directory entry, then: drop-in modules first, the script's directory second
    sys.path[:] = [pkg_dir, os.path.dirname(script)] + [p for p in sys.path if p not in (pkg_dir, os.path.dirname(script))]
    if _ROOT not in sys.path:
        sys.path.append(_ROOT)                      # the package's modules import their siblings through the package name
    sys.argv = [script] + list(argv)
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    if len(sys.argv) < 2:
        sys.exit("usage: python sivae_b200.py /path/to/soft_intro_vae/main.py [main.py arguments]")
    launch(sys.argv[1], sys.argv[2:])
else:
    _pkg = _package()
    sys.modules[__name__] = _pkg                    # `import sivae_b200` IS the package
