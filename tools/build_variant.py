"""Build an experimental variant of libinversus_b200.so with extra -D flags (kernel-shape
experiments; see profiles/). Usage: python tools/build_variant.py TAG -DINV_F32_T=256 ...
Writes profiles/variants/libinversus_b200.TAG.so; select it with INVERSUS_B200_LIB=<path>."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from inversus_b200 import _build  # noqa: E402

tag, flags = sys.argv[1], sys.argv[2:]
out = os.path.join(ROOT, "profiles", "variants", f"libinversus_b200.{tag}.so")
os.makedirs(os.path.dirname(out), exist_ok=True)
_build.build_library()  # makes sure host_expand.cpp.o exists
objs = [os.path.join(_build.HERE, os.path.basename(s) + ".o") for s in _build.HOST_SOURCES]
cmd = [_build.find_nvcc()] + _build.NVCC_FLAGS + flags + ["-o", out] + _build.SOURCES + objs
print(" ".join(cmd))
subprocess.check_call(cmd)
print(out)
