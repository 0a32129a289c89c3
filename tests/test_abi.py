"""C-ABI surface (no GPU needed): libamss_b200.so loads and exports every entry point declared in
include/amss.h, argument validation fails loudly with a message, and nothing in the shipped package
imports the oracle."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "adaptive-multispeaker-separation_b200")


def test_library_exports_every_declared_symbol():
    import amss_b200  # noqa: F401  (raises if the library is missing: there is no fallback)
    from amss_b200 import _lib
    decls = _lib.parse_header()
    assert len(decls) >= 40
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in decls if not hasattr(lib, n)]
    assert not missing, missing
    assert _lib.query("amss_version") >= 100


def test_header_cites_the_reference_interfaces_it_replaces():
    src = open(os.path.join(ROOT, "include", "amss.h")).read()
    for cite in ("models/adapt.py:95-134", "models/network.py:480-502", "utils/ops.py:358-383", "models/dpcl.py:41-86",
                 "models/Kmeans_2.py:14-188", "utils/ops.py:639-704"):
        assert cite in src, cite


def test_invalid_arguments_return_an_error_code_and_message():
    from amss_b200 import _lib
    rc = _lib.raw("amss_gemm")(None, 1, None, 1, None, 4, 4, 4, 0, 0, 0, 0, None, 4, 0, 0, None, 0, None)
    assert rc == -1 and "null pointer" in _lib.last_error()
    rc = _lib.raw("amss_filterbank_analysis_fwd")(None, None, 1, 8, 4, 4, 2, 2, 0, 0, None, None, None, 0, None)
    assert rc == -1


def test_size_queries_without_a_gpu():
    from amss_b200 import _lib
    assert _lib.query("amss_filterbank_analysis_out_frames", 64000, 1024, 256, 256, 0) == 250
    assert _lib.query("amss_blstm_saved_bytes", 4, 250, 256, 300) >= 2 * 250 * 4 * 5 * 300 * 4
    assert _lib.query("amss_gemm_workspace_bytes", 128, 128, 128, 0, 0, 1) > 2 * 128 * 128 * 2


def test_product_package_never_imports_the_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), os.path.join(dirpath, f)
