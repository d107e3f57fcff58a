"""CPU: the C-ABI shared library loads and exports exactly the symbols include/fastvim_b200.h
declares, and the ctypes signature table covers them (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fastvim_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fv_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built_lib():
    from fastvim_b200 import build

    return build.build()


def test_header_declares_entry_points():
    syms = declared_symbols()
    for s in ("fv_conv_pool_fwd", "fv_scan_fwd", "fv_gate_fwd", "fv_add_norm_fwd", "fv_selective_scan_fwd",
              "fv_last_error"):
        assert s in syms


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in the header but not exported"


def test_ctypes_table_matches_header(built_lib):
    from fastvim_b200 import _lib

    housekeeping = {"fv_last_error", "fv_version", "fv_launch_count", "fv_reset_launch_count"}
    assert set(_lib.SIGNATURES) == set(declared_symbols()) - housekeeping
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, args in _lib.SIGNATURES.items():
        decl = re.search(r"\b%s\s*\((.*?)\)\s*;" % name, src, flags=re.S).group(1)
        assert len([a for a in decl.split(",") if a.strip() and a.strip() != "void"]) == len(args), name
    l = _lib.lib()
    assert l.fv_version() >= 100
    assert l.fv_last_error() is not None


def test_argument_errors_are_reported_not_thrown(built_lib):
    """Errors come back as a status + message (no exceptions across the boundary)."""
    from fastvim_b200 import _lib

    l = _lib.lib()
    g = _lib.fv_geom(1, 6, 2, 2, 1, 2, 1, 0)  # dim not a multiple of 4
    rc = l.fv_conv_pool_fwd(ctypes.byref(g), 0, None, 8, 32, None, None, 1.0, 0, None, None)
    assert rc != 0 and b"dim" in l.fv_last_error()
    with pytest.raises(_lib.FastVimLibraryError):
        _lib.call("fv_add_norm_fwd", 7, 1, 4, None, 4, None, None, None, 1e-5, 1, None, 4, None, None, None, None)


def test_ops_refuse_cpu_tensors():
    import torch

    from fastvim_b200 import _lib, ops

    with pytest.raises(_lib.FastVimLibraryError):
        ops.conv_pool_fwd(torch.zeros(1, 4, 8), ops.Geometry.grid(2, 2), torch.zeros(2, 8, 4), None)


def test_compat_shims_resolve_reference_module_paths():
    """fastvim_b200/compat first on sys.path: the reference's own import lines (models/fastvim.py:9, 15-18;
    mamba_simple_faster.py:17-24) bind to the B200 implementation."""
    import importlib
    import sys

    compat = os.path.join(ROOT, "fastvim_b200", "compat")
    saved = {k: v for k, v in sys.modules.items() if k == "mamba_ssm" or k.startswith("mamba_ssm.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, compat)
    try:
        msf = importlib.import_module("mamba_ssm.modules.mamba_simple_faster")
        mscf = importlib.import_module("mamba_ssm.modules.mamba_simple_channel_faster")
        ln = importlib.import_module("mamba_ssm.ops.triton.layernorm")
        ssi = importlib.import_module("mamba_ssm.ops.selective_scan_interface")
        from fastvim_b200 import interface, mixer, mixer_channel, norm

        assert msf.Mamba is mixer.Mamba and mscf.Mamba is mixer_channel.Mamba
        assert ln.RMSNorm is norm.RMSNorm and ln.rms_norm_fn is norm.rms_norm_fn and ln.layer_norm_fn is norm.layer_norm_fn
        from fastvim_b200 import mixer_channel_2dcompress, mixer_masked
        assert importlib.import_module("mamba_ssm.modules.mamba_simple_masked_faster").Mamba_masked is mixer_masked.Mamba_masked
        assert importlib.import_module("mamba_ssm.modules.mamba_simple_masked_faster_v2").Mamba_masked is mixer_masked.Mamba_masked
        assert importlib.import_module("mamba_ssm.modules.mamba_simple_channel_faster_2dcompress").Mamba is mixer_channel_2dcompress.Mamba
        assert (importlib.import_module("faster_mamba_ssm.ops.selective_scan_interface").selective_scan_fn
                is interface.selective_scan_fn_compressed)
        for fn in ("selective_scan_fn", "mamba_inner_fn_no_out_proj", "mamba_inner_fn_no_out_proj_withoutZ",
                   "FastVim_mamba_inner_fn_no_out_proj_withoutZ"):
            assert getattr(ssi, fn) is getattr(interface, fn)
        # same parameter names / shapes as the reference module (state-dict compatible)
        m = msf.Mamba(32, token_size=[4, 6], layer_idx=0)
        names = set(dict(m.named_parameters()))
        assert {"in_proj.weight", "conv1d.weight", "conv1d_b.weight", "x_proj.weight", "x_proj_b.weight", "dt_proj.weight",
                "dt_proj_b.bias", "A_log", "A_b_log", "D", "D_b", "layernorm.weight", "out_proj.weight"} <= names
    finally:
        sys.path.remove(compat)
        for k in [k for k in sys.modules if k.split(".")[0] in ("mamba_ssm", "faster_mamba_ssm")]:
            del sys.modules[k]
        sys.modules.update(saved)
