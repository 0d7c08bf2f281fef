"""tests/golden/model_traces.json (tools/probe_model_trace.py, recorded from the real reference models) against the
figures SURVEY §8(a) quotes: call counts and separable forward GFLOP per image."""
import json
from pathlib import Path

import pytest

TRACES = json.loads((Path(__file__).parent / "golden" / "model_traces.json").read_text())


@pytest.mark.parametrize("name,calls,gflop,image", [("yolo11n", 87, 4.21, 1024), ("yolo11s", 87, 14.34, 1024),
                                                    ("qresnet34", 36, 1.85, 224)])
def test_trace_totals(name, calls, gflop, image):
    t = TRACES[name]
    rows = t["rows"]
    assert t["image"] == image and sum(r[8] for r in rows) == calls == t["calls"]
    total = sum(c * 4 * 2 * ho * ho * co * (ci // g) * k * k for ci, co, k, s, g, ho, bn, bias, c in rows) / 1e9
    assert abs(total - gflop) < 0.006 and abs(total - t["gflop_fwd_per_image"]) < 1e-9
    for ci, co, k, s, g, ho, bn, bias, c in rows:
        assert ci % g == 0 and co % g == 0 and k in (1, 3, 7) and s in (1, 2) and ho * s <= image


def test_yolo11n_iqbn_volume():
    # 84 IQBN calls, 41.0 M elements per image
    rows = TRACES["yolo11n"]["rows"]
    assert sum(r[8] for r in rows if r[6]) == 84
    assert abs(sum(r[8] * r[1] * r[5] * r[5] * 4 for r in rows if r[6]) / 1e6 - 41.0) < 0.05


def test_bench_loader_reads_the_fixture():
    import bench
    rows, meta = bench.load_model_trace("yolo11s")
    assert len(rows) == len(TRACES["yolo11s"]["rows"]) and meta["mix"] == "A"
