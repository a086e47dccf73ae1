"""TEST INFRASTRUCTURE ONLY -- import shim for the *unmodified* reference at /root/reference/MC-GRA.

Used only by tests/golden/make_golden.py inside the build container (the reference tree does not
exist on the GPU box).  It changes no hot-path arithmetic; it only makes the reference importable
on this image:
  * stub `matplotlib.pyplot` (only plt.cla() is called, MC-GRA/topology_attack.py:122) and
    `torchmetrics` (AUROC only used by the never-called metric(), topology_attack.py:15-21);
  * `numpy.int = int` (MC-GRA/utils.py:512-514 uses the alias removed in numpy >= 1.24).
"""
import sys
import types

import numpy as np

REF_DIR = "/root/reference/MC-GRA"


def install():
    if not hasattr(np, "int"):
        np.int = int  # noqa: NPY001 - the reference needs the removed alias
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        plt.cla = lambda *a, **k: None
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    if "torchmetrics" not in sys.modules:
        tm = types.ModuleType("torchmetrics")

        class AUROC:  # never constructed on the paths we run
            def __init__(self, *a, **k):
                raise NotImplementedError("torchmetrics stub")

        tm.AUROC = AUROC
        sys.modules["torchmetrics"] = tm
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)


def load():
    """Return the reference modules (topology_attack, utils, hsic, models.gcn, gcn_parameterized)."""
    install()
    import importlib
    mods = {}
    for name in ("utils", "base_attack", "topology_attack", "hsic", "models.gcn", "gcn_parameterized", "baseline"):
        mods[name] = importlib.import_module(name)
    return mods
