"""Loads tests/hostsim/libftl_hostsim.so: the device functions of csrc/*.cuh compiled for the CPU
(a TEST target so kernel logic can be checked in a container without a GPU; never a product path)."""
import os
import subprocess

from continiousenvironment_follower_leader_b200 import capi

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim")


def lib(name="libftl_hostsim.so"):
    subprocess.check_call(["make", "-C", _DIR, "-s", name])
    return capi.load(os.path.join(_DIR, name))


def make_env(gc, n, **kw):
    return capi.HostEnv(gc, n, lib=lib(), **kw)
