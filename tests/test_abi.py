"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU, exports exactly
what include/tess.h declares, and the product refuses to compute without a device (no fallback)."""
import ctypes
import importlib
import os
import sys
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "tess.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tess_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(tess):
    path = tess._lib.build()
    L = ctypes.CDLL(path)
    declared = _declared()
    assert len(declared) >= 30
    for s in declared:
        assert hasattr(L, s), f"{s} declared in include/tess.h but not exported"
    assert sorted(tess._lib.SYMBOLS) == declared
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (tess_[a-z0-9_]+)", out)))
    assert exported == declared, set(exported) ^ set(declared)


def test_library_is_sm100a_only(tess):
    out = subprocess.run(["cuobjdump", "-lelf", tess._lib.build()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_a_device(tess):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert tess.device_count() == 0
    with pytest.raises(tess.TessError) as e:
        tess.Diagram(0)
    assert e.value.code == -3  # TESS_ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under the package or include/ may reference it."""
    pkg = os.path.join(ROOT, "the-tessellator_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", "Makefile")):
                txt = open(os.path.join(base, f), errors="replace").read()
                assert "liboracle" not in txt and "oracle_binding" not in txt and "tess_oracle" not in txt, os.path.join(base, f)


def test_product_has_no_cpu_path():
    """The CPU warp emulator (tests/emu) is test infrastructure too.  The kernel sources carry its hooks behind
    TESS_WARP_EMU, which only tests/emu/shim defines: the package's Python never mentions it, the package's Makefile
    never defines it, and the built library holds no emulator symbol."""
    import subprocess

    pkg = os.path.join(ROOT, "the-tessellator_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            path = os.path.join(base, f)
            if f.endswith(".py") or f == "Makefile":
                txt = open(path, errors="replace").read()
                assert "TESS_WARP_EMU" not in txt and "warp_emu" not in txt and "libemu" not in txt and "tests/emu" not in txt, path
    lib = os.path.join(pkg, "libtess_b200.so")
    syms = subprocess.run(["nm", "-D", "-C", lib], capture_output=True, text=True).stdout + subprocess.run(["nm", "-C", lib], capture_output=True, text=True).stdout
    assert "emu::" not in syms and "emu_launch" not in syms


def test_generators_are_pure_integer_streams(gen):
    import numpy as np

    u = gen.uniform(4, 1)
    # splitmix64 known answers: mix(0) and the first draws of seed 1 (computed once, pinned here)
    assert int(gen.mix(np.array([0], dtype=np.uint64))[0]) == 0xE220A8397B1DCDAF
    assert u.shape == (4, 3) and np.all((u >= 0) & (u < 1))
    assert np.array_equal(gen.uniform(2, 1, start=2), u[2:])
    b = gen.bcc(4, 5)
    assert b.shape == (128, 3) and np.all((b > 0) & (b < 1))
    assert np.array_equal(gen.bcc(4, 5, start=10, count=7), b[10:17])


def test_rust_ffi_declares_only_functions_of_the_header():
    """rust/ cannot be compiled here (no rustc): at least every `fn tess_*` of ffi.rs must exist in include/tess.h with the
    same number of parameters, and every ffi function used by interface.rs must be declared."""
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "tess.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    c_decl = {m.group(1): m.group(2) for m in re.finditer(r"\b(tess_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr)}
    ffi = open(os.path.join(root, "rust", "src", "ffi.rs")).read()
    r_decl = {m.group(1): m.group(2) for m in re.finditer(r"pub fn (tess_[a-z0-9_]+)\s*\((.*?)\)\s*(?:->[^;]+)?;", ffi, flags=re.S)}
    assert len(r_decl) >= 30

    def arity(args: str) -> int:
        args = args.strip()
        return 0 if args in ("", "void") else args.count(",") + 1

    for name, args in r_decl.items():
        assert name in c_decl, f"{name} is not in tess.h"
        assert arity(args) == arity(c_decl[name]), f"{name}: {arity(args)} parameters in ffi.rs, {arity(c_decl[name])} in tess.h"

    # parameter TYPES, one by one: the C type of tess.h translated to the Rust type it must be bound as
    scalar = {"int": "c_int", "int32_t": "i32", "uint32_t": "u32", "int64_t": "i64", "uint64_t": "u64", "size_t": "usize", "double": "f64", "char": "c_char", "void": "c_void",
              "uint16_t": "u16", "tess_diagram": "tess_diagram", "tess_result": "tess_result", "tess_query": "tess_query", "tess_search": "tess_search",
              "tess_opts": "tess_opts", "tess_slab": "tess_slab"}

    def c_to_rust(decl: str) -> str:
        decl = decl.strip()
        decl = re.sub(r"\[[^\]]*\]\s*$", "*", re.sub(r"\b[A-Za-z_][A-Za-z0-9_]*\s*(\[[^\]]*\])?\s*$", lambda m: m.group(1) or "", decl, count=1).strip())  # drop the name; T x[6] -> T*
        toks = re.findall(r"const|\*|[A-Za-z_][A-Za-z0-9_]*", decl)
        toks = [t for t in toks if t != "struct"]
        base_const = toks[0] == "const"
        if base_const:
            toks = toks[1:]
        base, stars = toks[0], [t for t in toks[1:] if t in ("*", "const")]
        out = scalar[base]
        # pointers from the inside out: the innermost one points at a (const?) base
        const_next = base_const
        i = 0
        while i < len(stars):
            assert stars[i] == "*", decl
            out = ("*const " if const_next else "*mut ") + out
            const_next = i + 1 < len(stars) and stars[i + 1] == "const"
            i += 2 if const_next else 1
        return out

    def split_args(a: str):
        a = a.strip()
        return [] if a in ("", "void") else [x.strip() for x in a.split(",")]

    checked = 0
    for name, args in r_decl.items():
        want = [c_to_rust(a) for a in split_args(c_decl[name])]
        got = [re.sub(r"\s+", " ", a.split(":", 1)[1].strip()) for a in split_args(args)]
        assert got == want, f"{name}: ffi.rs binds {got}, tess.h declares {want}"
        checked += len(want)
    assert checked > 90
    # return types
    c_ret = {m.group(2): m.group(1).strip() for m in re.finditer(r"^\s*([A-Za-z_][A-Za-z0-9_ \*]*?)\s*\b(tess_[a-z0-9_]+)\s*\(", hdr, flags=re.M)}
    r_ret = {m.group(1): (m.group(2) or "").strip() for m in re.finditer(r"pub fn (tess_[a-z0-9_]+)\s*\(.*?\)\s*(?:->\s*([^;]+))?;", ffi, flags=re.S)}
    for name in r_decl:
        want = {"int": "c_int", "void": "", "uint64_t": "u64", "const char*": "*const c_char", "const char *": "*const c_char"}[c_ret[name]]
        assert r_ret[name] == want, f"{name}: returns {r_ret[name]!r} in ffi.rs, {c_ret[name]!r} in tess.h"
    used = set(re.findall(r"ffi::(tess_[a-z0-9_]+)", open(os.path.join(root, "rust", "src", "interface.rs")).read()))
    types = set(re.findall(r"pub struct (tess_[a-z0-9_]+)", ffi))
    assert used - types <= set(r_decl), used - types - set(r_decl)


def test_rust_repr_c_structs_match_the_header():
    """tess_opts and tess_slab cross the boundary by value/pointer: field order, names and types must match tess.h."""
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(root, "include", "tess.h")).read(), flags=re.S)
    ffi = open(os.path.join(root, "rust", "src", "ffi.rs")).read()
    ctype = {"double": "f64", "int64_t": "i64", "uint64_t": "u64", "uint32_t": "u32", "int32_t": "i32", "void*": "*mut c_void"}
    for name in ("tess_opts", "tess_slab"):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), hdr, flags=re.S).group(1)
        c_fields = []
        for decl in [d.strip() for d in body.split(";") if d.strip()]:
            m = re.match(r"(.*?)([A-Za-z_][A-Za-z0-9_]*)(\[(\d+)\])?$", decl)
            t = ctype[m.group(1).replace(" ", "")]
            c_fields.append((m.group(2), "[%s; %s]" % (t, m.group(4)) if m.group(4) else t))
        rbody = re.search(r"#\[repr\(C\)\]\s*(?:#\[derive\([^)]*\)\]\s*)?pub struct %s \{(.*?)\}" % name, ffi, flags=re.S).group(1)
        r_fields = [(m.group(1), m.group(2).strip()) for m in re.finditer(r"pub ([a-z_0-9]+): ([^,]+),", rbody)]
        assert r_fields == c_fields, (name, r_fields, c_fields)


def test_rust_build_script_compiles_the_makefiles_sources():
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    mk = open(os.path.join(root, "the-tessellator_b200", "csrc", "Makefile")).read()
    srcs = re.search(r"^SRCS\s*:=\s*(.*)$", mk, flags=re.M).group(1).split()
    rs = open(os.path.join(root, "rust", "build.rs")).read()
    listed = re.findall(r'"([a-z_]+\.cu)"', rs.split(".arg(\"-lcudart\")")[0])
    assert sorted(listed) == sorted(srcs), (listed, srcs)
    for flag in ("-fmad=false", "-prec-div=true", "-prec-sqrt=true", "arch=compute_100a,code=sm_100a"):
        assert flag in rs and flag in mk


def test_bench_refuses_an_ncu_capture_of_other_kernel_sources(tmp_path):
    """bench.py's `roofline` takes the instruction count per cell from profiles/clip_kernel_traffic.json, which carries a
    SHA-256 of the kernel sources it was captured from: a file whose hash does not match the sources is not used."""
    import importlib.util
    import json

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_for_test", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    sys.path.insert(0, os.path.join(root, "profiles"))
    from make_traffic_json import kernel_sources_sha256

    good = dict(warp_instructions_per_launch=1.8e11, cells_per_launch=10_000_000, dram_bytes_per_launch=6.2e9, issue_active_pct=85.8,
                kernel_sources_sha256=kernel_sources_sha256())
    pg, pb = tmp_path / "good.json", tmp_path / "bad.json"
    json.dump(good, open(pg, "w"))
    json.dump(dict(good, kernel_sources_sha256="0" * 64), open(pb, "w"))
    assert bench.ncu_traffic_per_launch(str(pg))["cells_per_launch"] == 10_000_000
    assert bench.ncu_traffic_per_launch(str(pb)) is None
    assert bench.ncu_traffic_per_launch(str(tmp_path / "missing.json")) is None
