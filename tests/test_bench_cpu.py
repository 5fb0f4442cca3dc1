"""CPU tests of the measurement code (bench.py): the algorithmic work figures of SURVEY 8(d), the roofline arithmetic,
the sharding plan the bench uses, and the reference arm on a tiny bounded sample (oracle/_ref/libref.so = the unmodified
netlib.cpp when it was built, else the numpy port).  No GPU."""
import importlib.util
import json
import os

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_pair_geometry_and_survey_work_figures(bench):
    """Config 2: FLOP = 10 P dD dM Nk Nl and bytes = 4 P (4 dD + 5 dM) per frame and pair; at 64 frames the three
    pairs are 374 GFLOP and 3.46 GB per step -- the figures the verdict's HBM fraction of the step is computed from."""
    w = bench.WORKLOADS["c2"]
    geo = bench.pair_geometry(w)
    assert geo == [(3, 16, 320, 240), (16, 32, 160, 120), (32, 64, 80, 60)]
    B, T = w["batch"], 25
    flop = sum(10.0 * nx * ny * dD * dM * T for dD, dM, nx, ny in geo) * B
    byts = sum(4.0 * nx * ny * (4 * dD + 5 * dM) for dD, dM, nx, ny in geo) * B
    assert abs(flop / 1e9 - 374) < 1.0
    assert abs(byts / 1e9 - 3.46) < 0.01
    assert bench.pair_geometry(bench.WORKLOADS["c3"]) == [(3, 16, 512, 512), (16, 32, 256, 256), (32, 64, 128, 128)]
    assert [g[:2] for g in bench.pair_geometry(bench.WORKLOADS["c4"])] == [(3, 16), (16, 32), (32, 64), (64, 128), (128, 256)]


def test_roofline_record_arithmetic(bench):
    pk = dict(hbm=6554.2, tf_burst=1600.0, tf_sust=1348.3, source="test")
    # a tensor-core kernel whose three BF16X3 passes outweigh its bytes: bound = tensor, frac = achieved / sustained peak
    top = dict(name="wgrad_ts", ms=1.5, launches=3, flops=3 * 49.8e9, bytes=3 * 0.54e9)
    r = bench.roofline_of(top, bench.WORKLOADS["c2"], "bf16x3", pk, 3.0, "c2")
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and r["mma_passes"] == 3
    assert np.isclose(r["achieved"], 49.8e9 / 0.5e-3 / 1e12)
    assert np.isclose(r["frac"], r["achieved"] / pk["tf_sust"]) and np.isclose(r["pipe_frac"], 3 * r["frac"])
    assert np.isclose(r["share_of_step"], 0.5) and np.isclose(r["avg_launch_ms"], 0.5)
    # a CUDA-core kernel is always read against HBM, on its algorithmic bytes
    top = dict(name="fft_rows_r2c", ms=0.92, launches=1, flops=1e10, bytes=2.4e9)
    r = bench.roofline_of(top, bench.WORKLOADS["c3"], "bf16x3", pk, 6.5, "c3")
    assert r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert np.isclose(r["achieved"], 2.4e9 / 0.92e-3 / 1e9) and np.isclose(r["frac"], r["achieved"] / pk["hbm"])
    assert r["algorithmic_bytes_per_launch"] == 2.4e9
    assert bench.roofline_of(None, bench.WORKLOADS["c2"], "bf16x3", pk, 1.0, "c2") is None


def test_committed_traffic_table_feeds_the_roofline(bench):
    """roofline.traffic comes from the committed ncu capture (profiles/traffic.json), per launch."""
    t = bench.ncu_traffic("c2", "wgrad_ts")
    assert t is not None and 4e8 < t < 6e8
    assert bench.ncu_traffic("c2", "no_such_kernel") is None


def test_bench_takes_frame_ownership_from_the_sharding_plan(bench):
    import dp
    assert bench.dp is dp
    owned = [dp.frame_range(r, 8, 64) for r in range(8)]
    assert [o[0] for o in owned] == [64 * r for r in range(8)] and all(o[1] == 64 for o in owned)


def test_reference_arm_on_a_tiny_sample(bench):
    """--impl reference in miniature: two parallel frame-samples of pair 0 on the smallest crop; the value is what was executed
    (frame-samples per second), the full-configuration projection is a separate record."""
    w = dict(bench.WORKLOADS["c2"])
    res = bench.cpu_reference_run(w, budget_s=0.05, cores=2, validate=False)
    assert res is not None and res["cores"] == 2 and res["kind"] in ("reference", "port")
    assert res["q"] == 8 and "40x30 crop" in res["sample"]
    assert res["value"] > 0 and np.isclose(res["value"], 2 / res["seconds"])
    assert "NOT full-configuration" in res["unit"]
    ex = res["extrapolated"]
    assert ex["value"] < res["value"] and ex["factor"] > 1
    json.dumps(res)  # the record goes into the JSON line as it is
