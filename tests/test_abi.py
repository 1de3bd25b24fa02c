"""CPU checks of the C-ABI boundary: the library builds for sm_100a, loads, and exports every
symbol include/summarizer_b200.h declares; the ctypes table and struct layout match the header.
No compute is launched (there is no GPU in the build container)."""
import ctypes
import os
import re

import numpy as np

from summarizer_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "summarizer_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(smz_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(N.LIB_PATH), "build with __graft_entry__.build()"
    L = ctypes.CDLL(N.LIB_PATH)
    names = _declared()
    assert len(names) >= 8
    for n in names:
        assert hasattr(L, n), f"{n} declared in the header but not exported"


def test_ctypes_table_covers_header():
    assert sorted(N.SIGNATURES) == _declared()


def test_desc_layout_matches_header():
    src = open(HEADER).read()
    body = src[src.index("typedef struct smz_video_desc"):src.index("} smz_video_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for typ, names in re.findall(r"(int64_t|int32_t)\s+([a-z0-9_,\s]+);", body):
        fields += [(n.strip(), typ) for n in names.split(",")]
    assert [f for f, _ in fields] == list(N.VIDEO_DESC.names)
    for (name, typ) in fields:
        assert N.VIDEO_DESC.fields[name][0].itemsize == (8 if typ == "int64_t" else 4)
    assert N.VIDEO_DESC.itemsize == 104


def test_version_and_error_strings_callable_without_gpu():
    L = N.lib()
    assert b"sm_100a" in L.smz_version()
    assert isinstance(L.smz_last_error(), bytes)


def test_no_oracle_import_in_product():
    """The product package must never import the oracle (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "summarizer_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
