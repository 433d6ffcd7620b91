"""The C-ABI shared library loads without a GPU and exports exactly what include/pixparse_b200.h declares; the ctypes
binding (prototypes AND argument-struct layouts) is checked against the header, field by field."""
import ctypes
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pixparse_b200.h")


def _header_source():
    src = open(HEADER).read()
    return re.sub(r"/\*.*?\*/", "", src, flags=re.S)


def _header_functions():
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", _header_source())))


def _ctype_of(decl):
    """C parameter / field declaration (without its name) -> ctypes type used by the binding."""
    from pixparse_b200 import _lib
    d = " ".join(decl.replace("*", " * ").split())
    m = re.match(r"const (B200\w+) \*$", d)
    if m:
        return ctypes.POINTER(_lib.STRUCTS[m.group(1)])
    if d.endswith("*"):
        return ctypes.c_void_p
    return {"int": ctypes.c_int, "long long": ctypes.c_longlong, "float": ctypes.c_float,
            "unsigned int": ctypes.c_uint}[d]


def _split_decl(param):
    """'const float* bias' -> ('const float*', 'bias')"""
    m = re.match(r"^(.*?[\s\*])([A-Za-z_]\w*)$", param.strip())
    assert m, param
    return m.group(1).strip(), m.group(2)


def _header_prototypes():
    """{name: (return C type, [param C types])} for every b200_* function the header declares."""
    src = re.sub(r"typedef struct.*?\}\s*\w+;", "", _header_source(), flags=re.S)
    src = re.sub(r"enum\s*\{.*?\};", "", src, flags=re.S)
    src = re.sub(r"^\s*#.*$", "", src, flags=re.M)            # preprocessor lines
    src = re.sub(r'extern\s+"C"\s*\{', "", src)
    out = {}
    for ret, name, params in re.findall(r"([\w\s\*]+?)\b(b200_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src):
        params = params.strip()
        plist = [] if params in ("", "void") else [_split_decl(p)[0] for p in params.split(",")]
        out[name] = (" ".join(ret.split()), plist)
    return out


def _header_structs():
    """{struct name: [(field C type, field name), ...]}"""
    out = {}
    for body, name in re.findall(r"typedef struct \w+ \{(.*?)\}\s*(\w+);", _header_source(), flags=re.S):
        fields = [_split_decl(f) for f in body.split(";") if f.strip()]
        out[name] = fields
    return out


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from pixparse_b200 import _lib
    lib = _lib.lib()
    names = _header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert lib.b200_abi_version() == _lib.ABI_VERSION == 3
    m = re.search(r"#define B200_ABI_VERSION (\d+)", open(HEADER).read())
    assert int(m.group(1)) == _lib.ABI_VERSION


def test_python_binding_table_matches_header():
    from pixparse_b200 import _lib
    declared = set(_header_functions())
    bound = set(_lib.SIGNATURES) | set(_lib.OTHER_SIGNATURES)
    assert declared == bound, (sorted(declared - bound), sorted(bound - declared))


def test_ctypes_argtypes_match_header_prototypes():
    """Every parameter of every prototype, in order: a transposed or mistyped ctypes signature fails here, not on the GPU."""
    from pixparse_b200 import _lib
    protos = _header_prototypes()
    assert set(protos) == set(_lib.SIGNATURES) | set(_lib.OTHER_SIGNATURES)
    for name, (ret, params) in protos.items():
        bound = _lib.SIGNATURES.get(name, _lib.OTHER_SIGNATURES.get(name))
        want = [_ctype_of(p) for p in params]
        assert bound == want, f"{name}: ctypes argtypes {bound} != header {want}"
        want_ret = {"int": ctypes.c_int, "long long": ctypes.c_longlong, "const char*": ctypes.c_char_p,
                    "const char *": ctypes.c_char_p}[ret]
        got_ret = _lib.RESTYPES.get(name, ctypes.c_int)
        assert got_ret == want_ret, f"{name}: restype {got_ret} != {want_ret}"
        fn = getattr(_lib.lib(), name)
        assert list(fn.argtypes or []) == want and fn.restype == want_ret     # what is really installed on the handle


def test_ctypes_struct_fields_match_header():
    from pixparse_b200 import _lib
    structs = _header_structs()
    assert set(structs) == set(_lib.STRUCTS)
    for sname, fields in structs.items():
        mirror = _lib.STRUCTS[sname]
        got = [(n, t) for n, t in mirror._fields_]
        want = [(n, _ctype_of(t)) for t, n in fields]
        assert got == want, f"{sname}: ctypes fields differ from the header:\n{got}\n{want}"
        assert fields[0] == ("unsigned int", "struct_size")


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_struct_layout_matches_the_c_compiler(tmp_path):
    """sizeof and every offsetof as gcc lays the header's structs out == what ctypes computes for the mirrors."""
    from pixparse_b200 import _lib
    structs = _header_structs()
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(void) {']
    for sname, fields in structs.items():
        lines.append(f'  printf("{sname} sizeof %zu\\n", sizeof({sname}));')
        for _, fname in fields:
            lines.append(f'  printf("{sname} {fname} %zu\\n", offsetof({sname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    n = 0
    for line in out.strip().splitlines():
        sname, fname, val = line.split()
        mirror = _lib.STRUCTS[sname]
        if fname == "sizeof":
            assert ctypes.sizeof(mirror) == int(val), (sname, ctypes.sizeof(mirror), val)
        else:
            assert getattr(mirror, fname).offset == int(val), (sname, fname)
        n += 1
    assert n > 100


def test_argument_errors_are_reported_without_a_gpu():
    """Shape / alignment / struct-version validation happens on the host before any CUDA call."""
    from pixparse_b200 import _lib
    lib = _lib.lib()
    rc = lib.b200_gemm_bf16(_lib.GemmArgs(lda=8, ldb=8, ldo=8), None)
    assert rc == -1
    assert b"empty problem" in lib.b200_last_error()
    bad = _lib.GemmArgs(m=128, n=128, k=64)
    bad.struct_size -= 8            # a caller compiled against an older header
    assert lib.b200_gemm_bf16(bad, None) == -1
    assert b"struct_size" in lib.b200_last_error()
    assert lib.b200_gemm_bf16(None, None) == -1
    rc = lib.b200_layernorm_fwd(_lib.LayerNormFwdArgs(rows=0, dim=768, eps=1e-5), None)
    assert rc == -1
    with pytest.raises(TypeError):
        _lib.GemmArgs(no_such_field=1)


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    """tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, TMA -> UTMALDG / UTMAREDG (B200_PROFILING.md)."""
    from pixparse_b200 import _lib
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMAREDG"):
        assert mnemonic in sass, mnemonic
    # the warp-level mma.sync path is confined to the 16-row weight-streaming linear of the greedy-decode step (HBM-bound
    # by > 8x, csrc/decode.cu); every contraction of the train step and of the teacher-forced decoder is tcgen05
    owner = None
    for line in sass.splitlines():
        if "Function :" in line:
            owner = line.split("Function :")[1].strip()
        elif "HMMA.16816" in line:
            assert owner is not None and "decode_linear_kernel" in owner, f"legacy mma.sync in {owner}"
