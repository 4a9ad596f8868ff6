"""The driver's contract with bench.py, checked without a GPU: the byte model behind `roofline` is SURVEY.md §8(d)'s, both arms
describe the workload identically, the committed bench lines of the end-of-round build carry every key the contract names and
are consistent with themselves, and the product arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _bench():
    import bench
    return bench


def test_byte_model_is_the_survey_formula():
    """bytes_view = 316 P + 172 R + 52 HW at SH degree 0; the stage split must add up to it in every variant
    (the fused variants only move bytes between stages), and the consumed-instance model charges the gather terms on R_consumed."""
    b = _bench()
    P, R, HW = 500_000, 7_113_985, 1200 * 680
    for fused_sort in (False, True):
        for fused_scatter in (False, True):
            st = b.roofline_bytes(P, R, HW, fused_sort, fused_scatter)
            assert sum(st.values()) == 316 * P + 172 * R + 52 * HW
            assert all(v >= 0 for v in st.values())
    st = b.roofline_bytes(P, R, HW)
    assert st["render_backward"] == 84 * R + 24 * HW + 40 * P and st["render_forward"] == 44 * R + 28 * HW
    assert st["preprocess"] == 108 * P and st["preprocess_backward"] == 168 * P and st["binning"] == 44 * R
    Rc = 974_202
    sc = b.roofline_bytes(P, R, HW, fused_sort=True, R_consumed=Rc)
    assert sc["render_backward"] == 84 * Rc + 24 * HW + 40 * P
    assert sc["render_forward"] == 44 * Rc + 28 * HW + 8 * R + 24 * Rc
    assert sc["render_forward"] < b.roofline_bytes(P, R, HW, fused_sort=True)["render_forward"]


def test_both_arms_describe_the_workload_identically():
    b = _bench()
    for name in ("C1_tum_tracking", "C2_replica_mapping", "C3_batched_tracking", "C4_large"):
        c = b.workload_config(name)
        assert set(c) == {"workload", "l2"} and c["workload"].startswith(name) and "model" not in c
    assert b.DEFAULT_WORKLOAD == "C2_replica_mapping"      # the configuration BASELINE.json's multi-GPU metric is quoted on
    assert "V=10" in b.workload_config("C2_replica_mapping")["workload"] and "V=32" in b.workload_config("C4_large")["workload"]


BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "clocks"}


def _load(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.load(f)


@pytest.mark.parametrize("name", ["r2s_bench_ours.json", "r2u_bench_ours_short.json"])
def test_committed_line_of_the_product_arm_keeps_the_contract(name):
    d = _load(name)
    assert BASE_KEYS | {"gpu_launches", "roofline"} <= set(d)
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["scaling"] == "strong" and d["config"]["workload"].startswith("C2_replica_mapping")
    assert d["value"] == pytest.approx(1e3 / d["ms_per_step"], rel=1e-6)
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e) and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]      # host copies and the host sync are inside the timed region: never the device-timed figure
    assert d["gpu_launches"] == d["steps"] * d["launches_per_step"]["kernels"] > 0
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], abs=2e-4) and 0 < r["frac"] <= 1.0
    assert r["achieved"] == pytest.approx(r["alg_bytes_per_launch"] / (r["kernel_ms"] * 1e-3) / 1e9, rel=2e-3)
    c = d["clocks"]
    assert c["sm_mhz"] > 0.9 * c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if "cpu_baseline" in d:
        assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"]) and d["cpu_baseline"]["kind"] in ("port", "reference")
    a = d["also_C1"]
    assert BASE_KEYS <= set(a) and a["scaling"] == "weak" and a["config"]["workload"].startswith("C1_tum_tracking")


def test_committed_line_of_the_reference_arm_keeps_the_contract():
    d, ours = _load("r2s_bench_ref.json"), _load("r2s_bench_ours.json")
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    for k in ("metric", "unit", "higher_is_better", "config"):      # the driver divides the two lines: same metric, same workload
        assert d[k] == ours[k], k
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert {"kind", "cores", "sample", "value"} <= set(d["cpu_baseline"]) and d["cpu_baseline"]["value"] == d["value"]
    # both arms print the pose gradient of the same last view: the parity the bench itself carries
    a, b = ours["check"]["dL_dtau_last"], d["check"]["dL_dtau_last"]
    scale = max(abs(x) for x in b)
    assert max(abs(x - y) for x, y in zip(a, b)) <= 1e-4 * scale


def test_product_arm_refuses_to_run_without_a_cuda_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout) and r.stdout.strip() == ""


def test_reference_arm_builds_its_inputs_without_the_product_library():
    """`--impl reference` must not map libgsr_b200.so.  Its input builders (shared with the product arm) are run
    in a fresh process whose memory map is then searched for the library."""
    code = (
        "import sys; sys.argv = ['bench.py']; sys.path.insert(0, %r)\n"
        "import bench\n"
        "bench.build_workload('C1_tum_tracking', 4, 0, 'cpu')\n"
        "bench.window_inputs('C3_batched_tracking', 2)\n"
        "bench.workload_config('C2_replica_mapping')\n"
        "print('MAPPED' if any('libgsr_b200' in l for l in open('/proc/self/maps')) else 'CLEAN')\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("CLEAN"), r.stderr[-2000:]
