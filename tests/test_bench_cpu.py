"""CPU legs of bench.py that need no GPU: the sequential reduced-algebra baseline (device code compiled for the host)."""
import sys

from tests.conftest import ROOT


def test_reduced_cpu_baseline_leg(pkg):
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import bench
    r = bench.reduced_cpu_rate(pkg, "cfg2")
    assert r["unit"] == bench.UNIT and r["cores"] == 1 and r["same_config"] is True and r["kind"] == "port"
    assert r["value"] > 0 and r["ms_per_iteration"] > 0
    assert r["gen_corrected"] > 0 and r["tight_rows"] > 0          # the correction path ran on the whole cfg2 case
    assert "1000 generators + 200 storages" in r["sample"]
