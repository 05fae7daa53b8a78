"""The product path must never import, link or execute anything under oracle/ (it is test infrastructure)."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "bcftools_b200")


def test_package_sources_do_not_reference_oracle():
    bad = []
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "oracle/" in txt or "libmcall_oracle" in txt or "libmcall_ref" in txt:
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_library_does_not_link_oracle():
    lib = os.path.join(PKG, "lib", "libmcall_b200.so")
    if not os.path.exists(lib):
        from bcftools_b200 import build
        build.build()
    out = subprocess.run(["ldd", lib], capture_output=True, text=True).stdout
    assert "oracle" not in out and "mcall_ref" not in out, out
    syms = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True).stdout
    assert "oracle_mcall_batch" not in syms and "ref_mcall_batch" not in syms
