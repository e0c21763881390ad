"""bench.py's host-side pieces that need no GPU: both arms describe the workload with the same `config`, the scaling
label follows the mode, the roofline denominator comes from MEASURED_PEAKS.json when the driver wrote one, and the NUMA
binding never raises."""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_config_is_a_function_of_the_workload_only():
    for wl in bench.WORKLOADS:
        cfg = bench.workload_config(wl)
        assert set(cfg) == {"workload", "l2"} and cfg["workload"].startswith(wl + ":")
        assert cfg == bench.workload_config(wl)                     # what --impl reference prints is the same dict
        json.dumps(cfg)
    assert "10000000 fingerprints x 120 hashes" in bench.workload_config("c3")["workload"]
    assert "5 file segments of 10000000 fingerprints" in bench.workload_config("c4")["workload"]
    assert bench.METRIC == json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"].split(";")[0]


def test_scaling_label():
    assert bench.scaling_of("c3", "replicated") == "weak"           # one batch per rank
    assert bench.scaling_of("c4", "replicated") == "strong"         # the 1 M-query batch is split over the ranks
    assert bench.scaling_of("c3", "sharded") == "strong"            # every rank sees the whole batch


def test_measured_peak_prefers_the_drivers_file():
    peak, src = bench.measured_peak()
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        assert peak == float(json.load(open(p))["hbm_gbs"]) and "MEASURED_PEAKS" in src
    else:
        assert peak == 6650.0 and "fallback" in src


def test_numa_binding_is_best_effort():
    before = os.sched_getaffinity(0)

    class Props:
        pci_domain_id, pci_bus_id, pci_device_id = 0xFFFF, 0xFF, 0x1F   # no such device in sysfs

    fake = types.SimpleNamespace(cuda=types.SimpleNamespace(get_device_properties=lambda dev: Props()))
    msg = bench.bind_near_gpu(fake, 0)
    assert isinstance(msg, str) and msg.startswith("not bound")
    assert os.sched_getaffinity(0) == before
    broken = types.SimpleNamespace(cuda=types.SimpleNamespace(get_device_properties=lambda dev: object()))
    assert bench.bind_near_gpu(broken, 0).startswith("not bound")
