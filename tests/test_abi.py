"""The C-ABI boundary: libtcar_b200.so loads on a GPU-less host and exports exactly what include/tcar_b200.h
declares (no compute calls here -- those are the `-m gpu` tests)."""
import ctypes
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tcar_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|long long)\s+(tcar_[a-z0-9_]+)\s*\(", src)))


def declared_arg_counts():
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    out = {}
    for name, args in re.findall(r"\b(?:int|long long)\s+(tcar_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        out[name] = len([a for a in args.split(",") if a.strip()])
    return out


def test_library_builds_and_exports_every_declared_symbol(native):
    import __graft_entry__ as ge
    ge.build()
    lib = ctypes.CDLL(native.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in tcar_b200.h but not exported"


def test_binding_table_matches_header(native):
    """ctypes signatures in _native.SIGNATURES cover every declared function with the right arity."""
    decl = declared_arg_counts()
    assert set(decl) == set(native.SIGNATURES), set(decl) ^ set(native.SIGNATURES)
    for name, n in decl.items():
        sig = native.SIGNATURES[name]
        # every asynchronous entry point takes the stream as its last argument; the size queries do not
        assert len(sig) == n, f"{name}: header has {n} args, binding has {len(sig)}"


def test_no_stray_exports(native):
    """Only the declared C symbols (plus toolchain internals) are exported with the tcar_ prefix."""
    out = subprocess.run(["nm", "-D", "--defined-only", native.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r"\sT\s+(tcar_[a-z0-9_]+)$", out, flags=re.M)))
    assert exported == declared_functions()


def test_sass_uses_blackwell_tensor_and_tma_paths(native):
    """The scoring kernels must be tcgen05 (UTC*MMA) + TMA (UTMALDG), not legacy mma.sync (HMMA)."""
    sass = subprocess.run(["cuobjdump", "-sass", native.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass or "UTCMMA" in sass or re.search(r"UTC\w*MMA", sass)
    assert "UTMALDG" in sass
    assert "LDTM" in sass
    assert not re.search(r"\bHMMA\b", sass)


def test_product_path_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "session-based-news-recommendation_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py") and fn != "smoke.py":
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f"{fn} imports the oracle"
