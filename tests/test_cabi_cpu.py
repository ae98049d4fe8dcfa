"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol
that include/phyml_b200.h declares, the ctypes binding knows all of them, and -- there being no CPU
fallback -- creating an instance on a machine without a CUDA device fails loudly."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "phyml_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(plk_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from phyml_b200 import engine

    lib = engine.load_library()
    syms = declared_symbols()
    assert len(syms) >= 28
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"declared in the header but not exported by libphyml_b200.so: {missing}"
    unbound = [s for s in syms if s not in engine.EXPORTS]
    assert not unbound, f"declared in the header but unknown to the ctypes binding: {unbound}"
    assert lib.plk_version().decode().startswith("phyml_b200")


def test_struct_layouts_match_the_header():
    from phyml_b200 import engine

    assert C.sizeof(engine._Config) == 8 * 4          # plk_config: 8 ints
    assert C.sizeof(engine._Side) == 8                # plk_side: {int tip; int clv;}
    assert C.sizeof(engine._Op) == 28                 # plk_op: dst, c1, pmat1, c2, pmat2
    assert engine.OP_DTYPE.itemsize == 28
    assert C.sizeof(engine._SprCand) == 32            # plk_spr_cand: {plk_side a; double l_a; plk_side b; double l_b;}
    import numpy as np
    assert engine.Engine._pars_ops([(1, 2, 3)]).dtype == np.int32 and engine.Engine._pars_ops([(1, 2, 3)]).nbytes == 12  # plk_pars_op


def test_shim_defines_every_re_bound_symbol():
    """integration/lk_b200_shim.c must define, with external linkage, every reference entry point INTEGRATION.md lists."""
    import subprocess

    obj = os.path.join(ROOT, "integration", "_build", "lk_b200_shim.o")
    if not os.path.exists(obj):
        pytest.skip("shim not built (needs /root/reference at build time)")
    out = subprocess.run(["nm", "--defined-only", obj], capture_output=True, text=True, check=True).stdout
    defined = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    for sym in ("Lk", "dLk", "Update_Partial_Lk", "Update_PMat_At_Given_Edge", "Update_Eigen_Lr", "Make_Tree_For_Lk",
                "Free_Tree_Lk", "Make_Edge_Lk", "posix_memalign", "aLRT", "Pars", "Pars_At_Given_Edge", "Update_Partial_Pars",
                "Make_Tree_For_Pars", "Free_Tree_Pars", "Ancestral_Sequences"):
        assert sym in defined, sym
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for sym in ("Update_Partial_Pars", "Pars_At_Given_Edge", "Ancestral_Sequences", "Make_Edge_Lk", "plk_spr_candidates"):
        assert sym in text, sym


def test_header_is_plain_c_and_cxx(tmp_path):
    """the boundary is a C ABI: include/phyml_b200.h must compile on its own as C99 and as C++ (no torch / CUDA types)"""
    import shutil
    import subprocess

    for cc, std, ext in (("gcc", "-std=c99", "c"), ("g++", "-std=c++11", "cpp")):
        if not shutil.which(cc):
            pytest.skip(f"{cc} not available")
        src = tmp_path / f"t.{ext}"
        src.write_text(f'#include "{HEADER}"\nint main(void) {{ plk_config c; plk_spr_cand s; plk_pars_op o; (void)c; (void)s; (void)o; '
                       "return sizeof(plk_op) == 28 ? 0 : 1; }\n")
        res = subprocess.run([cc, std, "-Wall", "-Werror", "-pedantic", "-fsyntax-only", str(src)], capture_output=True, text=True)
        assert res.returncode == 0, res.stderr
    includes = re.findall(r"#include\s*[<\"]([^>\"]+)[>\"]", open(HEADER).read())
    assert sorted(includes) == ["stddef.h", "stdint.h"], includes


def test_no_cpu_fallback():
    """Without a CUDA device plk_create must fail with PLK_ERR_CUDA and say why (never compute on the CPU)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from phyml_b200.engine import Engine, EngineError

    with pytest.raises(EngineError) as ei:
        Engine(4, 10, 4, 4, 10, 5)
    assert "no CUDA device" in str(ei.value) and "no CPU fallback" in str(ei.value)


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under phyml_b200/ or integration/ may reference it."""
    bad = []
    for sub in ("phyml_b200", "integration", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, sub)):
            for f in files:
                if f.endswith((".py", ".c", ".cu", ".cuh", ".h")):
                    txt = open(os.path.join(dirpath, f), errors="ignore").read()
                    if re.search(r"oracle_backend|plk_oracle|liboracle|import oracle|from oracle", txt):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad
