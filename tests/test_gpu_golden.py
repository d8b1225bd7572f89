"""CUDA path (through the C ABI) against the reference's golden outputs and the CPU oracle."""
import numpy as np
import pytest

from tests import common
from oracle import pyoracle

pytestmark = pytest.mark.gpu

_, _, _, EXP = common.load_chrY()


def _device_run(et, feats, hits, o, max_batch=1 << 22, **kw):
    from mmannot_b200 import device
    a = device.Annotator(et, strategy=o["strategy"], overlap=o["overlap"], rescue_threshold=o["rescue_threshold"],
                         read_stats=o["read_stats"], max_batch_hits=max_batch, **kw)
    try:
        a.load_features(feats)
        a.submit(0, hits)
        return a.finish(0)
    finally:
        a.close()


@pytest.mark.parametrize("case", sorted(EXP["cases"].keys()))
def test_device_matches_reference(case):
    from mmannot_b200 import device
    c = EXP["cases"][case]
    o = common.case_options(c["args"])
    et, feats, hitsF, _ = common.load_chrY(o["variant"])
    hits = common.restrand(hitsF, o["strand"])
    res = _device_run(et, feats, hits, o)
    table = common.table_of(et, device.values_by_mask(res["rows"]))
    assert table == c["table"]
    for k, v in c["stats"].items():
        assert res["stats"][k] == v, k


@pytest.mark.parametrize("batch", [1000, 4096, 45082])
@pytest.mark.parametrize("strategy", ["default", "ratio", "random"])
def test_device_batching_invariant(batch, strategy):
    """Small batches cut reads in the middle: the deferred path must give the same answer."""
    from mmannot_b200 import device
    et, feats, hitsF, _ = common.load_chrY()
    hits = common.restrand(hitsF, "U")
    o = {"strategy": strategy, "overlap": 1.0, "rescue_threshold": 1.0, "read_stats": False}
    ref = pyoracle.run(et.elem_line, et.elem_strand, et.elem_vicinity, feats, hits, strategy=strategy, overlap=1.0)
    res = _device_run(et, feats, hits, o, max_batch=batch)
    got = device.values_by_mask(res["rows"])
    assert set(got) == set(ref["rows"])
    for m, v in ref["rows"].items():
        assert abs(got[m] - v) <= 1e-9 * max(1.0, abs(v))
    for k, v in ref["stats"].items():
        assert res["stats"][k] == v, k


@pytest.mark.parametrize("shift", [3, 7, 12, 20])
def test_device_bin_width_invariant(shift):
    from mmannot_b200 import device
    et, feats, hitsF, _ = common.load_chrY()
    hits = common.restrand(hitsF, "F")
    o = {"strategy": "default", "overlap": 0.5, "rescue_threshold": 1.0, "read_stats": False}
    ref = pyoracle.run(et.elem_line, et.elem_strand, et.elem_vicinity, feats, hits, overlap=0.5)
    res = _device_run(et, feats, hits, o, bin_shift=shift)
    assert device.values_by_mask(res["rows"]) == ref["rows"]
    assert res["stats"] == ref["stats"]
