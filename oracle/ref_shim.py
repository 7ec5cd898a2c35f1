"""ORACLE SUPPORT — TEST INFRASTRUCTURE ONLY.

Imports the *untouched* reference modules from /root/reference.  The reference
files are flat but use package-relative imports (XceptionLSTMV.py:5,
``from .Xception import xception``), and ``xception(pretrained=True)`` downloads
weights over HTTP (Xception.py:31-34,211-212).  This shim:

  1. builds a scratch package ``<tmp>/RefModels`` of symlinks to the reference files,
  2. points TORCH_HOME at ``<tmp>/torch_home`` and pre-seeds
     ``hub/checkpoints/xception-43020ad28.pth`` with a seeded ``Xception().state_dict()``
     so the constructor works offline,
  3. imports the reference classes.

Nothing is copied into the repo and no reference file is modified.  The GPU box
has no /root/reference: ``available()`` is False there and callers must skip.
"""
from __future__ import annotations

import importlib
import os
import sys
import tempfile

REF_DIR = os.environ.get("XCP_REFERENCE_DIR", "/root/reference")
_FILES = ("Xception.py", "XceptionLSTMV.py", "XceptionLSTMA.py")
_state = {}


def available() -> bool:
    return all(os.path.isfile(os.path.join(REF_DIR, f)) for f in _FILES)


def load(seed: int = 1234):
    """Returns a namespace with Xception, xception, SeparableConv2d, Block,
    XceptionLSTMV, XceptionLSTMA taken from the reference."""
    if "ns" in _state:
        return _state["ns"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_DIR)
    import torch

    root = tempfile.mkdtemp(prefix="xcp_ref_")
    pkg = os.path.join(root, "RefModels")
    os.makedirs(pkg)
    open(os.path.join(pkg, "__init__.py"), "w").close()
    for f in _FILES:
        os.symlink(os.path.join(REF_DIR, f), os.path.join(pkg, f))
    sys.path.insert(0, root)
    xmod = importlib.import_module("RefModels.Xception")

    home = os.path.join(root, "torch_home")
    os.makedirs(os.path.join(home, "hub", "checkpoints"))
    os.environ["TORCH_HOME"] = home
    torch.hub.set_dir(os.path.join(home, "hub"))
    torch.manual_seed(seed)
    torch.save(xmod.Xception().state_dict(), os.path.join(home, "hub", "checkpoints", "xception-43020ad28.pth"))

    vmod = importlib.import_module("RefModels.XceptionLSTMV")
    amod = importlib.import_module("RefModels.XceptionLSTMA")

    class NS:
        pass

    ns = NS()
    ns.Xception = xmod.Xception
    ns.xception = xmod.xception
    ns.SeparableConv2d = xmod.SeparableConv2d
    ns.Block = xmod.Block
    ns.XceptionLSTMV = vmod.XceptionLSTMV
    ns.XceptionLSTMA = amod.XceptionLSTMA
    ns.root = root
    _state["ns"] = ns
    return ns
