"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/sss_b200.h declares, fails loudly without a GPU, and its tail-mode arithmetic agrees with
what the unmodified reference actually writes (oracle/_ref, when built)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import binding as B
from shared_simd_scan_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sss_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sss_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(L, name), f"libsss_b200.so lacks {name}"
    assert set(names) == set(_lib.SIGNATURES), "python signatures out of sync with the header"


def test_size_helpers_match_reference_formulas():
    L = _lib.lib()
    assert L.sss_packed_bytes(509, 9) == 829
    assert L.sss_unpack_bytes(509) == 2068
    assert L.sss_bitvec_bytes(12) == 34
    for w in (1, 9, 17, 32):
        for n in (0, 1, 7, 8, 1003, 1 << 34):
            assert L.sss_packed_bytes(n, w) == (n * w + 7) // 8 + 256


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for the CPU-only box")
    h = C.c_void_p()
    st = _lib.lib().sss_ctx_create(0, C.byref(h))
    assert st == -3 and "no CPU fallback" in _lib.error_string(st)
    from shared_simd_scan_b200 import api
    with pytest.raises(RuntimeError):
        api.Context.current()


def test_host_library_exports_dropin_symbols():
    so = os.path.join(ROOT, "shared_simd_scan_b200", "libsimd_scan_host.so")
    assert os.path.exists(so)
    import subprocess
    syms = subprocess.run(["nm", "-DC", "--defined-only", so], capture_output=True, text=True).stdout
    for name in ["sss_host::scan(", "sss_host::decompress(", "sss_host::shared_scan(", "sss_host::shared_scan_linear(",
                 "sss_host::compress(", "sss_host::scan_range(", "shared_scan_128_linear_simple_w("]:
        assert name in syms, name


@pytest.mark.skipif(not B.ref_available(9), reason="oracle/_ref not built")
@pytest.mark.parametrize("w", [1, 2, 3, 4, 7, 9, 13, 16])
def test_tail_modes_match_what_the_reference_writes(w):
    """Data without zeros, key 0: only padding can match, so the set bits and the last non-zero byte
    of the reference output reveal exactly how many codes it evaluated and how many bytes it wrote."""
    L = _lib.lib()
    for n in list(range(1, 80)) + [509, 1003]:
        codes = np.ones(n, dtype=np.uint32)
        packed = B.pack(codes, w)
        cases = [("128", _lib.TAIL_8), ("128_unrolled", _lib.TAIL_32), ("unvectorized", _lib.TAIL_CURSOR64)]
        for variant, mode in cases:
            ref, hits = B.ref_scan(variant, 0, packed, n, w)
            n_eval = L.sss_tail_eval_codes(n, w, mode)
            nbytes = L.sss_tail_out_bytes(n, w, mode)
            exp, ehits = B.scan_range(packed, n_eval, w, 0, 0, granule=1)
            full = np.zeros_like(ref)
            full[: len(exp)] = exp
            assert len(exp) <= nbytes <= len(ref)
            assert (ref == full).all() and hits == ehits, (w, n, variant, n_eval)
        ref = B.ref_shared_scan("128_parallel", [0], packed, n, w)[0]
        n_eval = L.sss_tail_eval_codes(n, w, _lib.TAIL_CURSOR32)
        exp, _ = B.scan_range(packed, n_eval, w, 0, 0, granule=1)
        full = np.zeros_like(ref)
        full[: len(exp)] = exp
        assert (ref == full).all(), (w, n, "parallel", n_eval)
