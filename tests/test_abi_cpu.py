"""The C-ABI library builds for sm_100a without a GPU, loads, and exports every symbol include/detrb.h declares;
the ctypes structures mirror the header's structs field for field.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "detrb.h")


def _header():
    return open(HEADER).read()


def test_library_builds_loads_and_exports_all_declared_symbols():
    from detr_tensorflow_b200 import _lib
    so = _lib.build()
    L = ctypes.CDLL(so)
    declared = set(re.findall(r"\b(detrb_[a-z0-9_]+)\s*\(", _header()))
    declared -= {"detrb_stream_t"}
    assert len(declared) >= 20
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/detrb.h but not exported by libdetrb.so"
    assert set(_lib.EXPORTS) <= declared
    L.detrb_version.restype = ctypes.c_int
    assert L.detrb_version() >= 100


@pytest.mark.parametrize("cname,pyname", [("detrb_igemm_t", "IgemmParams"), ("detrb_wgrad_t", "WgradParams"),
                                          ("detrb_attn_fwd_t", "AttnFwdParams"), ("detrb_attn_bwd_t", "AttnBwdParams")])
def test_ctypes_structs_mirror_header(cname, pyname):
    from detr_tensorflow_b200 import _lib
    h = _header()
    body = re.search(r"typedef struct \{([^}]*)\}\s*" + cname + ";", h).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for stmt in body.split(";"):
        stmt = stmt.strip()
        if not stmt:
            continue
        # "const detrb_bf16 *Q, *K, *V" / "int M, N, K" / "float *dW"
        m = re.match(r"(?:const\s+)?[A-Za-z_0-9]+\s+(.*)", stmt)
        for nm in m.group(1).split(","):
            fields.append(nm.strip().lstrip("*").strip())
    py = [f[0] for f in getattr(_lib, pyname)._fields_]
    assert py == fields, (py, fields)


def test_no_product_import_of_oracle():
    """the product path must not route through the oracle (or any CPU fallback)"""
    pkg = os.path.join(ROOT, "detr_tensorflow_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), os.path.join(dp, f)
                assert "cabi_emulator" not in src


def _prototypes():
    """function name -> number of parameters, from include/detrb.h"""
    h = re.sub(r"/\*.*?\*/", "", _header(), flags=re.S)
    out = {}
    for _, name, args in re.findall(r"\b(int|const char \*|void)\s*\*?\s*(detrb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", h):
        args = args.strip()
        out[name] = 0 if args in ("", "void") else len(args.split(","))
    return out


def test_every_call_site_passes_the_declared_number_of_arguments():
    """ctypes does not check arity: a call with a missing or extra argument would corrupt the callee's view of the stack.  Every
    `<lib>.detrb_xxx(...)` call in the package must pass exactly the parameters include/detrb.h declares, and the CPU emulator
    of the ABI (test infrastructure) must mirror the same signatures."""
    import ast
    import inspect
    protos = _prototypes()
    assert len(protos) >= 30
    pkg = os.path.join(ROOT, "detr_tensorflow_b200")
    seen = set()
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if not f.endswith(".py"):
                continue
            tree = ast.parse(open(os.path.join(dp, f)).read())
            for node in ast.walk(tree):
                if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr.startswith("detrb_"):
                    name = node.func.attr
                    assert name in protos, f"{f}: {name} is not declared in include/detrb.h"
                    assert not node.keywords and not any(isinstance(a, ast.Starred) for a in node.args), (f, name)
                    assert len(node.args) == protos[name], f"{f}:{node.lineno} {name}: {len(node.args)} arguments, header declares {protos[name]}"
                    seen.add(name)
    assert len(seen) >= 25, sorted(seen)
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cabi_emulator
    for name, fn in inspect.getmembers(cabi_emulator.FakeLib, inspect.isfunction):
        if name.startswith("detrb_") and name in protos:
            n = len(inspect.signature(fn).parameters) - 1          # minus self
            assert n == protos[name], f"cabi_emulator.FakeLib.{name}: {n} parameters, header declares {protos[name]}"


def test_handle_entry_points_fail_cleanly_without_a_gpu():
    """detrb_create on a box without a B200 returns an error code and a message (no fallback, no crash); the handle functions
    reject NULL; detrb_destroy(NULL) is a no-op.  (The per-thread switch behaviour is a -m gpu test: tests/test_api_gpu.py.)"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    from detr_tensorflow_b200 import _lib
    L = ctypes.CDLL(_lib.build())
    L.detrb_last_error.restype = ctypes.c_char_p
    h = ctypes.c_void_p()
    rc = L.detrb_create(ctypes.c_int(0), ctypes.byref(h))
    assert rc < 0 and not h.value and L.detrb_last_error()
    assert L.detrb_create(ctypes.c_int(0), None) == -1
    assert L.detrb_destroy(None) == 0
    v = ctypes.c_int()
    assert L.detrb_handle_set(None, ctypes.c_int(0), ctypes.c_int(1)) == -1
    assert L.detrb_handle_get(None, ctypes.c_int(0), ctypes.byref(v)) == -1
    assert L.detrb_bind(None) == -1 and b"live handle" in L.detrb_last_error()
    # the option table of the Python wrapper mirrors the header's enum
    from detr_tensorflow_b200 import ops
    enum = dict((k.lower(), int(n)) for k, n in re.findall(r"DETRB_OPT_([A-Z_]+) = (\d+)", _header()))
    count = enum.pop("count")
    assert ops.OPTIONS == enum and count == len(enum)
