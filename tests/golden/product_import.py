"""Import a module of the PRODUCT's reference-named package tree (text-detection-ctpn_b200/lib/..., ctpn/...) in a process
that has the reference's own `lib` package imported (the golden generators and container-only fuzz scripts)."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = os.path.join(ROOT, "text-detection-ctpn_b200")


def load_product_module(name):
    """Returns the product's module `name` (e.g. 'lib.text_connector.detectors'); sys.modules' reference entries are put back."""
    saved = {k: sys.modules.pop(k) for k in [k for k in sys.modules if k == "lib" or k.startswith("lib.")]}
    sys.path.insert(0, PKG)
    try:
        mod = importlib.import_module(name)
        assert os.path.realpath(mod.__file__).startswith(PKG), mod.__file__
        return mod
    finally:
        sys.path.remove(PKG)
        for k in [k for k in sys.modules if k == "lib" or k.startswith("lib.")]:
            sys.modules.pop(k)
        sys.modules.update(saved)
