"""The C-ABI libraries load on a box without a GPU and export every symbol include/*.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from tests import common

ROOT = common.ROOT


def declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set()
    for m in re.finditer(r"\b(mm[ah]_[a-z_0-9]+)\s*\(", text):
        names.add(m.group(1))
    return sorted(names)


@pytest.fixture(scope="module", autouse=True)
def _build():
    common.ensure_built(("host", "cuda"))


def test_device_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(os.path.join(ROOT, "mmannot_b200", "lib", "libmmannot_b200.so"))
    names = declared("mmannot_b200.h")
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    from mmannot_b200 import device
    assert sorted(device.EXPORTS) == names


def test_host_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(os.path.join(ROOT, "mmannot_b200", "lib", "libmmannot_host.so"))
    names = declared("mmannot_b200_host.h")
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n


def test_no_cpu_fallback_without_a_device():
    """mma_create must fail loudly (MMA_ERR_NO_DEVICE) where there is no GPU; with one it must succeed."""
    import numpy as np
    from mmannot_b200 import device, host
    et = host.ElementTable(np.zeros(2, np.uint16), np.zeros(2, np.uint8), np.zeros(2, np.uint8))
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        has_gpu = False
    if has_gpu:
        device.Annotator(et).close()
    else:
        with pytest.raises(device.MmaError) as e:
            device.Annotator(et)
        assert e.value.code == -5


def test_version_string():
    from mmannot_b200 import device
    assert b"sm_100a" in device.lib().mma_version()
    assert device.lib().mma_dominant_kernel() in (b"k_batch_lean", b"k_batch_fast")


def test_every_exported_symbol_is_documented():
    """INTEGRATION.md section 5 lists every symbol of the device ABI with the reference interface it replaces."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [n for n in declared("mmannot_b200.h") if "`" + n + "`" not in doc]
    assert not missing, missing
