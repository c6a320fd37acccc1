"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/a3vlm_b200.h declares; no compute call is made (no GPU here)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from a3vlm_b200 import _lib

    assert os.path.exists(_lib.LIB_PATH), "build with python -c 'import __graft_entry__ as g; g.build()'"
    lib = _lib.load(strict=True)
    assert not _lib.MISSING
    text = open(_lib.HEADER).read()
    declared = set(re.findall(r"\b(a3_\w+)\s*\(", re.sub(r"/\*.*?\*/", "", text, flags=re.S)))
    assert declared == set(_lib.PROTOS), declared ^ set(_lib.PROTOS)
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.a3_version() >= 100
    assert isinstance(lib.a3_launch_count(), int)


def test_signatures_have_no_torch_types():
    from a3vlm_b200 import _lib

    allowed = {"int", "float", "long long", "void", "const void*", "void*", "const float*", "float*",
               "const long long*", "long long*", "const int*", "int*", "const char*", "unsigned char*"}
    for name, (ret, args) in _lib.PROTOS.items():
        assert ret.replace(" ", "") in {a.replace(" ", "") for a in allowed}, (name, ret)
        for t, _ in args:
            assert t.replace(" ", "") in {a.replace(" ", "") for a in allowed}, (name, t)


def test_ops_refuse_cpu_tensors_and_model_refuses_cpu_only_hosts():
    from a3vlm_b200 import ops
    from a3vlm_b200._lib import A3Error

    with pytest.raises(A3Error):
        ops.rmsnorm_fwd(torch.zeros(2, 8, dtype=torch.bfloat16), torch.ones(8, dtype=torch.bfloat16), 1e-5)
    if not torch.cuda.is_available():
        from a3vlm_b200.model import MetaModel

        with pytest.raises(A3Error):
            MetaModel("llama_b200", dict(dim=256, n_layers=1, n_heads=2, vocab_size=64, multiple_of=64), None)


def test_sass_is_blackwell_native():
    """The shipped binary carries tcgen05 MMAs (UTC*MMA), TMEM loads (LDTM) and TMA (UTMALDG) and no
    legacy HMMA tensor path -- checked from the cubin, no GPU needed."""
    import shutil
    import subprocess

    from a3vlm_b200 import _lib

    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "LDTM" in sass and "UTMALDG" in sass and "UTMAREDG" in sass
    assert " HMMA" not in sass


def test_product_code_never_imports_the_oracle_or_a_cpu_fallback():
    """The oracle is test infrastructure: nothing under a3vlm_b200/ may import, load or execute it (only tests/,
    __graft_entry__.smoke() and bench.py's CPU-baseline / reference arm do), and no product module may import the
    reference tree."""
    import ast
    import os

    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "a3vlm_b200")
    offenders = []
    for dirpath, _, files in os.walk(root):
        for fn in files:
            if not fn.endswith(".py"):
                continue
            path = os.path.join(dirpath, fn)
            tree = ast.parse(open(path).read())
            for node in ast.walk(tree):
                names = []
                if isinstance(node, ast.Import):
                    names = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom):
                    names = [node.module or ""]
                elif isinstance(node, ast.Constant) and isinstance(node.value, str) and node.value.startswith("/root/reference"):
                    names = [node.value]
                for n in names:
                    if "oracle" in n or n.startswith("accessory") or n.startswith("/root/reference"):
                        offenders.append((os.path.relpath(path, root), n))
    assert not offenders, offenders


def test_every_python_call_site_passes_the_declared_number_of_arguments():
    """The wrappers bind by position through ctypes: a call with the wrong number of arguments only fails on a GPU
    box.  Statically compare every `<lib>.a3_xxx(...)` call in the package (and bench.py) with the header."""
    import ast

    from a3vlm_b200 import _lib

    files = [os.path.join(ROOT, "bench.py")]
    for d, _, names in os.walk(os.path.join(ROOT, "a3vlm_b200")):
        files += [os.path.join(d, n) for n in names if n.endswith(".py")]
    seen = set()
    for path in files:
        for node in ast.walk(ast.parse(open(path).read())):
            if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr.startswith("a3_"):
                name = node.func.attr
                assert name in _lib.PROTOS, (path, node.lineno, name)
                if any(isinstance(a, ast.Starred) for a in node.args):
                    continue
                assert not node.keywords, (path, node.lineno, name)
                want = len(_lib.PROTOS[name][1])
                assert len(node.args) == want, f"{os.path.relpath(path, ROOT)}:{node.lineno}: {name} takes {want} arguments, call passes {len(node.args)}"
                seen.add(name)
    # every compute entry point is reachable from Python
    unused = set(_lib.PROTOS) - seen - {"a3_last_error", "a3_version"}
    assert not unused, unused
