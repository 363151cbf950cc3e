"""Loads tests/hostsim/libftl_hostsim.so: the device functions of csrc/*.cuh compiled for the CPU
(a TEST target so kernel logic can be checked in a container without a GPU; never a product path)."""
import os
import subprocess

from continiousenvironment_follower_leader_b200 import capi

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim")


# the host build only has the host-buffer entry points; it is bound here, by the tests, and never by the package
_HOST_ENTRY_POINTS = ("ftl_abi_version", "ftl_last_error", "ftl_create", "ftl_destroy", "ftl_rays_per_env", "ftl_num_envs",
                      "ftl_upload_scenarios", "ftl_reset_host", "ftl_step_host", "ftl_step_host_ex", "ftl_get_state",
                      "ftl_set_state")
_LIBS = {}


def lib(name="libftl_hostsim.so"):
    subprocess.check_call(["make", "-C", _DIR, "-s", name])
    path = os.path.join(_DIR, name)
    if path not in _LIBS:
        import ctypes
        L = ctypes.CDLL(path)
        names = [n for n in capi.SIGNATURES if hasattr(L, n)]
        assert set(_HOST_ENTRY_POINTS) <= set(names), "host build lacks %s" % (set(_HOST_ENTRY_POINTS) - set(names))
        _LIBS[path] = capi.bind(L, names)
    return _LIBS[path]


def make_env(gc, n, **kw):
    return capi.HostEnv(gc, n, lib=lib(), **kw)
