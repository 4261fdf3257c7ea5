"""The drop-in boundary: libnalu_b200.so loads without a GPU and exports every entry point
include/nalu_b200.h declares (and nothing in the Python binding's list is missing from the header).
No compute call is made here."""
import ctypes
import os
import re

from nalu_b200 import build as nb
from nalu_b200 import lib as nl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    with open(os.path.join(ROOT, "include", "nalu_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)  # comments mention names that are not declarations
    return sorted(set(re.findall(r"\b(nalu_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_entry_point():
    so = nb.build()
    L = ctypes.CDLL(so)
    names = _declared()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in include/nalu_b200.h but not exported: {missing}"


def test_python_binding_lists_the_header():
    assert sorted(nl.SYMBOLS) == _declared()


def test_version_and_error_string_without_a_device():
    L = nl.lib()
    assert L.nalu_b200_version() > 0
    assert isinstance(L.nalu_b200_last_error(), (bytes, type(None)))
