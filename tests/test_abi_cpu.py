"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports
every symbol include/vectorlite_cuda.h declares, and fails loudly (no CPU fallback) without a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def vl():
    import __graft_entry__ as g
    g.build()
    import vectorlite_b200
    return vectorlite_b200


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "vectorlite_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vl_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(vl):
    L = vl.lib()
    syms = _declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/vectorlite_cuda.h but not exported"
    # and the Python binding binds exactly the declared set
    assert sorted(L._vl_signatures) == syms


def test_no_torch_types_in_abi():
    src = open(os.path.join(ROOT, "include", "vectorlite_cuda.h")).read()
    assert "torch" not in src and "at::" not in src and "#include <cuda" not in src


def test_version_and_error_string(vl):
    L = vl.lib()
    assert b"sm_100a" in L.vl_version()
    assert isinstance(L.vl_last_error(), bytes)


def test_fails_loudly_without_gpu(vl):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(vl.VectorLiteError) as e:
        vl.FlatIndex(3)
    assert e.value.code == vl.VL_ERR_CUDA and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "vectorlite_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "vl_oracle" not in txt, f


def test_invalid_arguments(vl):
    L = vl.lib()
    h = C.c_void_p()
    assert L.vl_flat_create(0, 0, C.byref(h)) == vl.VL_ERR_INVALID      # dim == 0
    assert L.vl_flat_create(3, 0, None) == vl.VL_ERR_INVALID
    assert L.vl_index_len(None) == 0
    L.vl_index_destroy(None)                                             # no-op, must not crash


def test_shard_group_argument_checks(vl):
    """vl_group_* (the single-process shard group) rejects null / empty input without touching a device."""
    L = vl.lib()
    g = C.c_void_p()
    assert L.vl_group_create(None, 0, C.byref(g)) == vl.VL_ERR_INVALID and not g.value
    assert b"at least one" in L.vl_last_error()
    arr = (C.c_void_p * 1)(None)
    assert L.vl_group_create(arr, 1, C.byref(g)) == vl.VL_ERR_INVALID
    assert L.vl_group_create(arr, 1, None) == vl.VL_ERR_INVALID
    assert L.vl_group_size(None) == 0
    L.vl_group_destroy(None)                                             # no-op, must not crash


def test_bench_measured_arm_uses_oracle_only_as_checker():
    """bench.py: the only code of the measured arm (`main`) that names `oracle` is the cpu_baseline leg — the
    inputs come from the product's generator, the reference arm lives in run_reference()."""
    import ast
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    main = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "main")
    leg = next(n for n in ast.walk(main) if isinstance(n, ast.If) and "no_cpu_baseline" in ast.unparse(n.test))
    inside = {id(n) for n in ast.walk(leg)}
    stray = [n.lineno for n in ast.walk(main)
             if ((isinstance(n, ast.Name) and n.id == "oracle") or
                 (isinstance(n, ast.Import) and any(a.name == "oracle" for a in n.names))) and id(n) not in inside]
    assert not stray, f"bench.py main() touches oracle outside the cpu_baseline leg at lines {stray}"


def test_rust_bindings_name_only_exported_symbols(vl):
    """rust/vectorlite-cuda-sys/src/lib.rs (the extern "C" crate of INTEGRATION.md; cannot be compiled here) binds
    only functions the header declares and the library exports, with the declared number of arguments."""
    L = vl.lib()
    hdr = open(os.path.join(ROOT, "include", "vectorlite_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = {m.group(1): m.group(2) for m in re.finditer(r"\b(vl_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S)}
    rs = open(os.path.join(ROOT, "rust", "vectorlite-cuda-sys", "src", "lib.rs")).read()
    rs = re.sub(r"//.*", "", rs)
    bound = {m.group(1): m.group(2) for m in re.finditer(r"pub fn (vl_[a-z0-9_]+)\s*\(([^)]*)\)", rs, flags=re.S)}
    assert len(bound) >= 18

    def nargs(a):
        a = a.strip()
        return 0 if a in ("", "void") else a.count(",") + 1

    for name, args in bound.items():
        assert name in declared, f"{name} is bound in Rust but not declared in the header"
        assert hasattr(L, name), f"{name} is bound in Rust but not exported"
        assert nargs(args) == nargs(declared[name]), (name, args, declared[name])
    # the shim in INTEGRATION.md / rust/vectorlite-src-index/cuda.rs calls only what the sys crate binds
    shim = open(os.path.join(ROOT, "rust", "vectorlite-src-index", "cuda.rs")).read()
    for name in set(re.findall(r"sys::(vl_[a-z0-9_]+)\s*\(", shim)):
        assert name in bound, f"cuda.rs calls sys::{name}, which the sys crate does not bind"


def test_prepared_patches_still_apply():
    """experiments/*.patch (compile-checked changes waiting for GPU validation) apply to the current tree."""
    import glob
    import subprocess
    for patch in sorted(glob.glob(os.path.join(ROOT, "experiments", "*.patch"))):
        r = subprocess.run(["git", "apply", "--check", patch], cwd=ROOT, capture_output=True, text=True)
        assert r.returncode == 0, (patch, r.stderr)


def test_adversarial_rounding_counter_example_arithmetic():
    """experiments/adversarial_bf16_rounding.py: the arithmetic of the input on which a constant 0.0040 bound with BOTH
    operands rounded certifies a wrong top-k while the measured-norm bound refuses (DESIGN.md §3 caveat)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("vl_adv", os.path.join(ROOT, "experiments", "adversarial_bf16_rounding.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    m.cpu_demo()
