"""CPU oracle for the Shared_SIMD_Scan hot path -- TEST INFRASTRUCTURE ONLY.

oracle.c / oracle.h : plain-C restatement of the reference algorithm (parity PINNED, see oracle.h)
ref_shim.cpp        : C doorway into the unmodified reference, built per bit width by build_ref.sh
binding.py          : ctypes access to both for tests/, smoke() and bench.py's CPU legs

Nothing in shared_simd_scan_b200/ imports this package.
"""
