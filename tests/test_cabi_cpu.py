"""CPU-side checks of the boundary: the shared library loads, exports every symbol the header
declares, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes
import os
import subprocess

import pytest

from votca_b200 import _capi


def test_header_parses():
    protos = _capi.parse_header()
    assert len(protos) >= 55
    for must in ("gwbse_ctx_create", "gwbse_mmn_fill_block", "gwbse_mmn_mul_right", "gwbse_rpa_epsilon",
                 "gwbse_sigma_x", "gwbse_sigma_ppm_eval", "gwbse_bse_matmul", "gwbse_gramschmidt_dev"):
        assert must in protos


def test_library_exports_every_declared_symbol():
    api = _capi.capi()
    out = subprocess.run(["nm", "-D", "--defined-only", _capi.LIBPATH], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if ln.strip()}
    for name in api.protos:
        assert name in exported, name
        assert getattr(api.lib, name) is not None


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from votca_b200.api import Context, GwbseError
    with pytest.raises(GwbseError, match="needs a CUDA device"):
        Context(0)
