"""GPU: short runs of the randomised soaks (tests/soak/stress_*.py, fixed seeds) -- random thresholds, counting-filter
lengths, batch cuts, slab geometries and shapes against the oracle / zlib.  compute-sanitizer is closed on this pool, so
randomised parity is what stands in for racecheck; the long runs of the same scripts found the stale-extent bug fixed in
bloom_count.cuh (test_gpu_bloom.py::test_staged_records_end_in_an_overhang)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("script,budget,seed", [("stress_levels.py", 20, 101), ("stress_other.py", 15, 102), ("stress_medium.py", 25, 103),
                                                ("stress_thresholds.py", 20, 104), ("stress_first_touch.py", 30, 105), ("stress_search_exit.py", 20, 106)])
def test_randomised_soak(script, budget, seed):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "soak", script), str(budget), str(seed)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "stress ok" in r.stdout
