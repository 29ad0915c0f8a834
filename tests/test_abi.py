"""CPU: the C-ABI library builds, loads and exports exactly what include/p2t_b200.h declares; the
host layer refuses to run without CUDA (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "p2t_b200.h")


def _header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(p2t_\w+)\s*\(", src)))


@pytest.fixture(scope="module")
def built(p2t):
    import __graft_entry__ as entry
    entry.build()
    return p2t


def test_header_declares_the_expected_surface():
    fns = _header_functions()
    for must in ("p2t_gemm_bf16", "p2t_adapter_fwd", "p2t_adapter_bwd", "p2t_pool_fwd", "p2t_l2norm_fwd",
                 "p2t_infonce_ce", "p2t_similarity", "p2t_rows_plan", "p2t_last_error"):
        assert must in fns


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built._lib.LIB_PATH)
    for name in _header_functions():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", built._lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (p2t_\w+)", out)))
    assert exported == _header_functions(), "exported symbols and header declarations differ"


def test_ctypes_table_covers_the_header(built):
    bound = set(built._lib.SIGNATURES) | set(built._lib.NON_STATUS)
    assert bound == set(_header_functions())
    assert built._lib.load().p2t_abi_version() == 2


def test_argument_errors_are_reported_without_a_gpu(built):
    # null pointers are rejected before any CUDA call, with a message
    with pytest.raises(built.P2TError, match="null pointer"):
        built._lib.call("p2t_l2norm_fwd", None, 1, 8, None, None, None, None)
    with pytest.raises(built.P2TError, match="mask_bytes"):
        built._lib.call("p2t_rows_plan", 1, 3, 1, 1, 64, 1, 1, 1, 1, None, None, None)


def test_sass_is_blackwell_native(built):
    """tcgen05.mma / TMA / tcgen05.ld must be present in the shipped binary (UTCHMMA, UTMALDG, LDTM)."""
    sass = subprocess.run(["cuobjdump", "-sass", built._lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic
    assert "HMMA.16816" not in sass  # no legacy mma.sync path
    # the peer-memory exchange publishes / polls its flags at system scope (csrc/peer.cu)
    for mnemonic in ("MEMBAR.ALL.SYS", "STG.E.STRONG.SYS", "LDG.E.STRONG.SYS"):
        assert mnemonic in sass, mnemonic


def test_no_cpu_fallback(built):
    ad = built.ModalityAdapter(built.ModalityAdapterConfig(input_dim=16, intermediate_dim=32, output_dim=24))
    with pytest.raises(built.P2TError, match="CUDA"):
        ad(torch.randn(2, 3, 16))
    with pytest.raises(built.P2TError, match="CUDA"):
        built.BatchInfoNCELoss()(torch.randn(4, 16), torch.randn(4, 16))
    with pytest.raises(built.P2TError):
        built.readout_embeddings(torch.randn(2, 3, 8), torch.ones(2, 3, dtype=torch.long), "mix")


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the product package may import, load or execute it."""
    pkg_dir = os.path.join(ROOT, "prot2text-v2-esm3_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "oracle." not in text and "oracle/" not in text, f


def test_oracle_is_only_reached_from_the_checker_sites():
    """Outside tests/, the oracle may be imported by bench.py's CPU-baseline function and by smoke() only."""
    for f in os.listdir(os.path.join(ROOT, "tools")):
        if f.endswith((".py", ".sh")):
            assert "oracle" not in open(os.path.join(ROOT, "tools", f)).read(), f
    bench = open(os.path.join(ROOT, "bench.py")).read()
    sites = [m.start() for m in re.finditer(r"^\s*from oracle\b|^\s*import oracle\b", bench, flags=re.M)]
    assert 1 <= len(sites) <= 2  # the reference loader and the restatement, both inside the CPU-baseline function
    for site in sites:
        fn_start = bench.rfind("\ndef ", 0, site)
        assert bench[fn_start:].lstrip().startswith("def cpu_step_time")
    entry = open(os.path.join(ROOT, "__graft_entry__.py")).read()
    sites = [m.start() for m in re.finditer(r"^\s*from oracle\b|^\s*import oracle\b", entry, flags=re.M)]
    assert len(sites) == 1 and entry[entry.rfind("\ndef ", 0, sites[0]):].lstrip().startswith("def smoke")
    # build() may BUILD the checker side (oracle/build_ref.py packs the reference's hot-path files) but never runs it
    assert "build_ref.py" in entry[entry.find("def build"):entry.find("def smoke")]
