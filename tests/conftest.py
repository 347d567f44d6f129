"""pytest configuration: registers the `gpu` marker and puts the repo root and
the product directory (`text-detection-ctpn_b200/`, whose name is not an
importable identifier) on sys.path."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "text-detection-ctpn_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on the B200 box)")
