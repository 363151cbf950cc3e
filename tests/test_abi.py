"""CPU checks of the boundary: the ctypes mirror matches the header, the CUDA library loads here
(no compute calls without a GPU) and exports every symbol include/ftl.h declares."""
import ctypes as C
import os
import re

import pytest

from continiousenvironment_follower_leader_b200 import abi, build, capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "ftl.h")).read()
    return sorted(set(re.findall(r"\b(ftl_[a-z_]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    syms = _declared_symbols()
    for s in ("ftl_create", "ftl_destroy", "ftl_upload_scenarios", "ftl_reset", "ftl_step", "ftl_step_host",
              "ftl_reset_host", "ftl_get_state", "ftl_set_state", "ftl_stats", "ftl_last_error"):
        assert s in syms


def test_cuda_library_builds_loads_and_exports_every_symbol():
    out, _ = build.build()
    lib = C.CDLL(out)
    for s in _declared_symbols():
        assert hasattr(lib, s), "libftl.so does not export " + s
    lib.ftl_abi_version.restype = C.c_int
    assert lib.ftl_abi_version() == abi.FTL_ABI_VERSION


def test_struct_sizes_match_the_compiled_oracle():
    import oracle_py
    L = oracle_py.lib()
    assert L.ftl_oracle_sizeof_env_state() == C.sizeof(abi.FtlEnvState)
    assert L.ftl_oracle_sizeof_config() == C.sizeof(abi.FtlConfig)


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(capi.FtlLibraryMissing):
        capi.load(str(tmp_path / "libftl.so"))


def test_invalid_configuration_is_rejected_without_a_gpu():
    from continiousenvironment_follower_leader_b200.config import GameConfig
    lib = capi.load()
    gc = GameConfig()
    gc.c.corridor_cap = 100          # not a power of two
    h = C.c_void_p()
    rc = lib.ftl_create(C.byref(gc.c), 4, 0, 0, C.byref(h))
    assert rc == abi.FTL_ERR_INVALID and b"corridor_cap" in lib.ftl_last_error()


def test_product_loader_refuses_a_library_without_the_device_entry_points():
    """The host build of the device functions (tests/hostsim) lacks ftl_step / ftl_stats ...: only the tests' own loader
    (tests/hostsim_py.py) binds it; capi.load must not accept it as a stand-in for libftl.so."""
    import hostsim_py
    hostsim_py.lib()
    with pytest.raises(capi.FtlLibraryMissing):
        capi.load(os.path.join(ROOT, "tests", "hostsim", "libftl_hostsim.so"))


def test_every_declared_entry_point_has_a_ctypes_signature():
    assert set(_declared_symbols()) - {"ftl_set_error_message"} <= set(capi.SIGNATURES)
