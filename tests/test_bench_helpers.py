"""bench.py's host-side helpers (no GPU): the committed ncu summary is readable and the
roofline / issue figures taken from it serialise."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_ncu_summary_helpers():
    import bench

    traffic, src = bench.ncu_traffic_bytes("k_run_compare")
    assert src and src.startswith("ncu_full_summary_r") and 1e7 < traffic < 1e9
    issue = bench.ncu_issue_counters("k_run_compare")
    assert issue["source"] == src
    assert 0 < issue["issue_active_pct"] <= 100 and 0 < issue["fp64_pipe_pct"] <= 100
    assert issue["warp_instructions"] > 1e6
    assert bench.ncu_issue_counters("no_such_kernel") is None
    assert bench.ncu_traffic_bytes("no_such_kernel")[0] is None
    us, lst = bench.ncu_launch_list_us("k_weights_scan")
    assert lst.startswith("ncu_launches_r") and 1.0 < us < 100.0
    assert 20.0 < bench.ncu_launch_list_us("k_run_compare")[0] < 200.0
    assert bench.ncu_launch_list_us("no_such_kernel")[0] is None
    json.dumps({"issue": issue, "traffic": traffic})
    peak, kind = bench.measured_peak_gbs()
    assert peak > 1000 and kind in ("measured", "fallback")
    cfg = bench.config_dict(2)
    assert "workload" in cfg and "model" not in cfg and cfg["parallelism"] == "chains x2"
