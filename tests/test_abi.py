"""CPU tests of the boundary: the C-ABI library loads and exports every symbol include/p2g.h declares, the header and
the ctypes structs agree, and the product never routes through the oracle."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "p2g.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(p2g_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    import plonky2_goldibear_b200 as p2g
    lib = p2g.load_library()  # ImportError if the .so has not been built
    declared = _declared_symbols()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), "missing export " + name
    from plonky2_goldibear_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared
    assert lib.p2g_abi_version() == 1


def test_struct_layout_matches_header(tmp_path):
    """sizeof/offsetof of the two descriptor structs as the C compiler sees them == the ctypes mirrors."""
    from plonky2_goldibear_b200 import _lib
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "p2g.h"\nint main(){'
                   'printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(p2g_commit_desc), offsetof(p2g_commit_desc, col_ptrs),'
                   'offsetof(p2g_commit_desc, leaf_pitch), offsetof(p2g_commit_desc, shard_index),'
                   'offsetof(p2g_commit_desc, stage_seconds), sizeof(p2g_merkle_desc), offsetof(p2g_merkle_desc, leaves),'
                   'offsetof(p2g_merkle_desc, seconds)); return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    C, M = _lib.CommitDesc, _lib.MerkleDesc
    want = [ctypes.sizeof(C), C.col_ptrs.offset, C.leaf_pitch.offset, C.shard_index.offset, C.stage_seconds.offset,
            ctypes.sizeof(M), M.leaves.offset, M.seconds.offset]
    assert got == want


def test_no_device_fails_loudly():
    """Without a GPU p2g_init must fail with NO_DEVICE -- never fall back to a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import plonky2_goldibear_b200 as p2g
    with pytest.raises(p2g.P2GError) as e:
        p2g.Context(0)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "plonky2_goldibear_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f)).read()
                assert "p2o" not in text and "oracle/" not in text and "import oracle" not in text, os.path.join(dirpath, f)
    out = subprocess.run(["ldd", os.path.join(pkg, "libp2g_cuda.so")], capture_output=True, text=True).stdout
    assert "libp2o" not in out


def test_flag_constants_match_header():
    """Every P2G_FLAG_* the header defines has the same value in the ctypes mirror (the Rust p2g-sys file follows the same
    list, integration/rust/p2g_sys.rs)."""
    from plonky2_goldibear_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "p2g.h")).read()
    flags = dict(re.findall(r"#define\s+P2G_(FLAG_[A-Z_]+)\s+(\d+)u", hdr))
    assert len(flags) >= 6
    for name, value in flags.items():
        assert getattr(_lib, name) == int(value), name
    vals = sorted(int(v) for v in flags.values())
    assert vals == [1 << i for i in range(len(vals))], "flags must stay distinct bits"
    rust = open(os.path.join(ROOT, "integration", "rust", "p2g_sys.rs")).read()
    for name, value in flags.items():
        assert re.search(r"pub const P2G_%s: u32 = %s;" % (name, value), rust), "p2g_sys.rs lacks P2G_" + name


def test_rust_bindings_are_generated_from_the_header():
    """integration/rust/p2g_sys.rs is what tools/gen_rust_bindings.py makes of include/p2g.h today: every exported symbol
    declared, both #[repr(C)] structs with the header's field order, every constant."""
    res = subprocess.run(["python", os.path.join(ROOT, "tools", "gen_rust_bindings.py"), "--check"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    rust = open(os.path.join(ROOT, "integration", "rust", "p2g_sys.rs")).read()
    for name in _declared_symbols():
        assert re.search(r"pub fn %s\(" % name, rust), "p2g_sys.rs lacks " + name
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "p2g.h")).read(), flags=re.S)
    bodies = {name: body for body, name in re.findall(r"typedef struct \{([^}]*)\}\s*(p2g_[a-z_]+);", hdr)}
    for struct in ("p2g_commit_desc", "p2g_merkle_desc"):
        c_fields = []
        for stmt in bodies[struct].split(";"):
            for decl in stmt.split(","):
                ids = re.findall(r"[A-Za-z_][A-Za-z0-9_]*", re.sub(r"\[\d+\]", "", decl))
                if ids:
                    c_fields.append(ids[-1])  # the declarator's name is its last identifier
        r_body = re.search(r"pub struct %s \{(.*?)\n\}" % struct, rust, flags=re.S).group(1)
        r_fields = ["repr" if f == "repr_" else f for f in re.findall(r"pub ([a-z_0-9]+):", r_body)]
        assert r_fields == c_fields, (struct, r_fields, c_fields)


def test_rust_shim_has_no_elided_bodies():
    """cuda_shim.rs and dump_vectors.rs are complete files (not compiled here, but nothing is left to the reader)."""
    for name in ("cuda_shim.rs", "dump_vectors.rs"):
        text = open(os.path.join(ROOT, "integration", "rust", name)).read()
        code = "\n".join(ln.split("//")[0] for ln in text.splitlines())
        assert "/*" not in code and "todo!" not in code and "unimplemented!" not in code and " .. }" not in code, name
        assert code.count("{") == code.count("}") and code.count("(") == code.count(")"), name
    shim = open(os.path.join(ROOT, "integration", "rust", "cuda_shim.rs")).read()
    for needle in ("pub fn from_values(", "pub fn from_coeffs(", "pub fn new(leaves: Vec<Vec<F>>, cap_height: usize)",
                   "p2g_commit(ctx, &mut desc)", "p2g_merkle(ctx, &mut desc)", "Polynomial degrees inconsistent"):
        assert needle in shim, needle
