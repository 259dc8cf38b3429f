"""CPU: the C-ABI library loads and exports every symbol include/msda_b200.h declares.
No compute calls (there is no GPU here)."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "msda_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(msda_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for s in ("msda_forward_f32", "msda_forward_f64", "msda_forward_bf16", "msda_backward_f32",
              "msda_backward_f64", "msda_backward_bf16", "msda_debug_corners_f32",
              "msda_backward_workspace_bytes", "msda_has_fast_path", "msda_abi_version", "msda_build_info",
              "msda_last_error", "msda_launch_count"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(str(ROOT / "richsem_b200" / "lib" / "libmsda_b200.so"))
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/msda_b200.h but not exported"


def test_binding_loads_and_reports():
    from richsem_b200 import _capi

    assert _capi.lib.msda_abi_version() == _capi.MSDA_ABI_VERSION
    assert "sm_100a" in _capi.build_info()
    assert ctypes.sizeof(_capi.MsdaOpts) == 56  # struct msda_opts layout on LP64
    assert _capi.lib.msda_has_fast_path(4, 32, 4, 4) == 1
    assert _capi.lib.msda_has_fast_path(8, 32, 4, 4) == 0
    assert _capi.lib.msda_has_fast_path(4, 64, 4, 4) == 0
    # workspace query is pure host arithmetic
    assert _capi.lib.msda_backward_workspace_bytes(2, 22223, 8, 32, 4, 22223, 4) > 4 * 4 * 2 * 22223 * 8 * 16


def test_product_package_does_not_import_the_oracle():
    pkg = ROOT / "richsem_b200"
    for py in pkg.rglob("*.py"):
        src = py.read_text()
        assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("# oracle", ""), py
