"""The C-ABI library: loads on a CPU-only box, exports every prototype of include/b200serve.h,
and refuses to compute without a GPU (no CPU fallback)."""
import os
import re

import pytest

from clearml_serving_b200 import build, native
from tests.conftest import HAS_GPU, ROOT


def _header_prototypes():
    text = open(os.path.join(ROOT, "include", "b200serve.h")).read()
    return sorted(set(re.findall(r"B2S_API[^;(]*?\b(b2s_\w+)\s*\(", text)))


def test_library_builds_and_exports_every_header_symbol():
    build.build()
    lib = native.lib()
    names = _header_prototypes()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "libb200serve.so does not export {}".format(n)
    bound = sorted(p[0] for p in native.PROTOTYPES)
    assert bound == names, "native.PROTOTYPES and include/b200serve.h disagree"
    assert lib.b2s_abi_version() == 2


def test_sass_is_sm100a():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", native.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


@pytest.mark.skipif(HAS_GPU, reason="checks the no-GPU behaviour")
def test_no_cpu_fallback_without_gpu():
    with pytest.raises(native.B2SError) as ei:
        native.init(0)
    assert "no CPU fallback" in str(ei.value)
    with pytest.raises(native.B2SError):
        native.Model(native.MODEL_FOREST, b"\0" * 128, device=0)


def test_missing_library_is_loud(monkeypatch, tmp_path):
    monkeypatch.setattr(native, "_lib", None)
    monkeypatch.setattr(native, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        native.lib()
