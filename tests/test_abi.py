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


def test_argument_validation_needs_no_gpu():
    """Entry points reject bad arguments with SMZ_ERR_ARG (< 0) and a message before touching the device."""
    from summarizer_b200.models.lstm_stack import LstmDecode, LstmSeq
    L = N.lib()
    seq = (LstmSeq * 1)()
    seq[0].T, seq[0].H, seq[0].B = 4, 1000, 1                      # hidden size the recurrence does not implement
    assert L.smz_lstm_seq_forward(seq, 1, None, None) < 0 and b"hidden size" in L.smz_last_error()
    seq[0].H, seq[0].B = 1024, 9                                   # more sequences than one launch takes
    assert L.smz_lstm_seq_forward(seq, 1, None, None) < 0 and b"sequences per launch" in L.smz_last_error()
    assert L.smz_lstm_seq_forward(seq, 3, None, None) < 0          # one or two directions
    dec = LstmDecode()
    dec.T, dec.H, dec.B = 0, 2048, 1
    assert L.smz_lstm_decode_forward(ctypes.byref(dec), None, None) < 0 and b"empty sequence" in L.smz_last_error()
    nbytes = ctypes.c_int64(0)
    assert L.smz_kts_workspace_bytes(100, 5, 1, ctypes.byref(nbytes)) == 0 and nbytes.value > 100 * 100 * 4
    one = ctypes.c_void_p(16)
    assert L.smz_kts(one, None, 10, 0, 0, 20, 1, 100, 0, 1.0, 1, one, one, one, one, 1 << 30, None) < 0
    assert b"(m+1)*lmin" in L.smz_last_error()
    assert L.smz_cvt_bf16_multi(None, None, None, 9, None) < 0 and b"1..8 segments" in L.smz_last_error()
    assert L.smz_host_pack_user_summary(None, 3, None, None, None, 1) < 0


def test_struct_sizes_match_header():
    """ctypes mirrors of the LSTM structs against the sizes the C compiler gives the header's."""
    import subprocess, tempfile
    from summarizer_b200.models.lstm_stack import LstmDecode, LstmSeq
    from summarizer_b200.models.dsn import DsnParams
    from summarizer_b200.models.vasnet import VasnetParams
    from summarizer_b200.models.vasnet_autograd import VasnetGrads
    src = '#include <stdio.h>\n#include "summarizer_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(smz_lstm_seq), sizeof(smz_lstm_decode), sizeof(smz_dsn_params), sizeof(smz_video_desc), sizeof(smz_vasnet_params), sizeof(smz_vasnet_grads));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", os.path.join(d, "s")])
        sizes = [int(x) for x in subprocess.check_output([os.path.join(d, "s")]).split()]
    assert sizes == [ctypes.sizeof(LstmSeq), ctypes.sizeof(LstmDecode), ctypes.sizeof(DsnParams), N.VIDEO_DESC.itemsize,
                     ctypes.sizeof(VasnetParams), ctypes.sizeof(VasnetGrads)]


def test_episode_sampling_argument_validation():
    L = N.lib()
    one = ctypes.c_void_p(16)
    assert L.smz_bernoulli_logprob(None, 10, 5, one, None, one, one, None) < 0 and b"NULL pointer" in L.smz_last_error()
    assert L.smz_bernoulli_logprob(one, 0, 5, one, None, one, one, None) < 0
    assert L.smz_bernoulli_logprob(one, 10, 0, one, None, one, one, None) < 0 and b"episodes" in L.smz_last_error()
    assert L.smz_bernoulli_logprob_backward(one, one, None, 10, 5, one, None) < 0
