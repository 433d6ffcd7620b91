"""The C-ABI shared library loads without a GPU and exports exactly what include/pixparse_b200.h declares."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "pixparse_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from pixparse_b200 import _lib
    lib = _lib.lib()
    names = _header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert lib.b200_abi_version() == 1


def test_python_binding_table_matches_header():
    from pixparse_b200 import _lib
    declared = set(_header_functions())
    bound = set(_lib.SIGNATURES) | {"b200_last_error", "b200_attention_bwd_workspace_bytes",
                                     "b200_preprocess_workspace_bytes"}
    assert declared == bound, (sorted(declared - bound), sorted(bound - declared))


def test_argument_errors_are_reported_without_a_gpu():
    """Shape / alignment validation happens on the host before any CUDA call."""
    from pixparse_b200 import _lib
    lib = _lib.lib()
    rc = lib.b200_gemm_bf16(None, 8, 0, None, 8, 0, 0, 0, 0, 0, None, 8, None, 0, None, None, 0, 0, 0, None)
    assert rc == -1
    assert b"empty problem" in lib.b200_last_error()
    rc = lib.b200_layernorm_fwd(None, None, None, None, None, None, None, 0, 768, 1e-5, None)
    assert rc == -1


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    """tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, TMA -> UTMALDG / UTMAREDG (B200_PROFILING.md)."""
    import shutil
    import subprocess
    from pixparse_b200 import _lib
    if shutil.which("cuobjdump") is None:
        import pytest
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMAREDG"):
        assert mnemonic in sass, mnemonic
    assert "HMMA.16816" not in sass      # no legacy mma.sync tensor path
