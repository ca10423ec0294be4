"""The C-ABI library loads without a GPU and exports every function include/tc_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "tc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"static inline[^{]*\{.*?\n\}\n", "", text, flags=re.S)      # header-only helpers (tc_te_decode, tc_seq_apply_delta)
    return sorted(set(re.findall(r"\b(tc_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_declared_symbols():
    import __graft_entry__ as g
    g.build()
    from transcriptclean_b200 import engine
    lib = ctypes.CDLL(engine.LIB_PATH)
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(engine.EXPORTS) == names


def test_no_gpu_is_a_loud_error():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from transcriptclean_b200 import engine
    with pytest.raises(engine.EngineError):
        engine.Engine(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "transcriptclean_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no CPU fallback", ""), f
                assert "hostsim" not in src.lower() or f in ("tc_b200.cu", "tc_plan.cuh"), f


def test_header_is_plain_c():
    """include/tc_b200.h compiles as C99 (what a cgo / JNI / ctypes-free C consumer would include) and the
    record layouts are the ones engine.py mirrors."""
    import subprocess
    import tempfile
    from transcriptclean_b200 import engine
    src = ('#include "tc_b200.h"\n#include <stdio.h>\n'
           'int main(void){ tc_te r; uint64_t w[2] = {((uint64_t)3 << 56) | ((uint64_t)1 << 58) | ((uint64_t)4 << 59) | ((uint64_t)7 << 32) | 1000u, 2000u};\n'
           ' const uint8_t in[6] = {65, 67, 71, 84, 65, 65}, dl[6] = {(0 << 6) | 2, (1 << 6) | 1, (2 << 6) | 2, 110, 110, (0 << 6) | 3}; uint8_t out[8];\n'
           ' if (tc_te_decode(w, &r) != 2 || r.a != 1000 || r.b != 2000 || r.size != 7 || r.type != 3 || r.status != 1 || r.reason != 4) return 1;\n'
           ' if (tc_seq_apply_delta(in, 6, dl, 6, out, 8) != 7 || out[0] != 65 || out[1] != 67 || out[2] != 110 || out[4] != 84 || out[6] != 65) return 2;\n'
           ' if (tc_seq_apply_delta(in, 6, dl, -1, out, 8) != 6 || tc_seq_apply_delta(in, 6, dl, 6, out, 5) != -1) return 3;\n'
           ' printf("%zu %zu %zu %zu %zu\\n", sizeof(tc_te), sizeof(tc_batch_in), sizeof(tc_batch_out), '
           'sizeof(tc_options), sizeof(tc_timing)); return 0; }\n')
    d = tempfile.mkdtemp()
    open(d + "/t.c", "w").write(src)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           "-o", d + "/t", d + "/t.c"])
    sizes = [int(x) for x in subprocess.check_output([d + "/t"]).split()]
    assert sizes == [engine.TE_DTYPE.itemsize, ctypes.sizeof(engine._BatchIn), ctypes.sizeof(engine._BatchOut),
                     ctypes.sizeof(engine._Options), ctypes.sizeof(engine.Timing)]


def test_host_only_entry_points_work_without_a_gpu():
    """tc_cig_pack16 (16-bit CIGAR words + listed ops) and tc_device_count are host code: callable on a machine without a GPU."""
    import numpy as np
    from transcriptclean_b200 import engine
    lib = engine.load_library()
    assert lib.tc_device_count() >= 0
    op = np.frombuffer(b"MIDNSHP=XBMQM", dtype=np.uint8).copy()
    ln = np.array([5, 1, 4095, 4096, 7, 2, 3, 9, 8, 6, 0, 4, -1], dtype=np.int32)
    w = np.zeros(len(op) + 8, dtype=np.uint16)
    ei, eo, el = np.zeros(8, np.int64), np.zeros(8, np.uint8), np.zeros(8, np.int32)
    m = lib.tc_cig_pack16(op.ctypes.data, ln.ctypes.data, len(op), w.ctypes.data, 8, ei.ctypes.data, eo.ctypes.data, el.ctypes.data, 1)
    assert m == 4 and list(ei[:4]) == [3, 10, 11, 12]            # 4096N (too long), 0M, the letter Q, -1M
    assert [chr(c) for c in eo[:4]] == ["N", "M", "Q", "M"] and list(el[:4]) == [4096, 0, 4, -1]
    assert all(w[k] == 0xFFFF for k in (3, 10, 11, 12))
    for k in (0, 1, 2, 4, 5, 6, 7, 8, 9):
        assert "MIDNSHP=XB"[w[k] & 15] == chr(op[k]) and (w[k] >> 4) == ln[k]
    assert lib.tc_cig_pack16(op.ctypes.data, ln.ctypes.data, len(op), w.ctypes.data, 3, ei.ctypes.data, eo.ctypes.data, el.ctypes.data, 1) == -1
