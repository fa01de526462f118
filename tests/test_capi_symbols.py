"""The C-ABI library loads on a machine without a GPU, exports every symbol include/opencorr_b200.h
declares, and refuses to run (loudly) when there is no CUDA device -- there is no CPU fallback."""
import os
import re

import pytest

from opencorr_b200 import _capi
import conftest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "opencorr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ocb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    lib = _capi.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libopencorr_b200.so does not export %s" % n
    assert sorted(_capi.SIGNATURES) == names, "opencorr_b200/_capi.py and include/opencorr_b200.h disagree"


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under opencorr_b200/ or include/ may import, link or call it."""
    banned = ("import oracle", "from oracle", "liboc_oracle", "oco_", "oc_oracle")
    for top in ("opencorr_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    text = open(os.path.join(dirpath, f)).read()
                    for b in banned:
                        assert b not in text, "%s/%s references the oracle (%r)" % (top, f, b)


@pytest.mark.skipif(conftest.HAVE_GPU, reason="a CUDA device is present")
def test_no_gpu_means_loud_failure():
    import opencorr_b200 as ob
    with pytest.raises(ob.OpenCorrB200Error) as ei:
        ob.Engine(0)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)
