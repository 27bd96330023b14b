"""CPU-only: the caller combiner (csrc/combiner.h — concurrent single-query callers → one batched search) against a
stand-in search, compiled with g++.  Every caller gets the answer to its own query, failures stay with their caller,
callers of different (k, metric, ef) classes never share a batch, nothing hangs — with polling waiters and without."""
import json
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def stress_bin(tmp_path_factory):
    if not shutil.which("g++"):
        pytest.skip("needs g++")
    out = str(tmp_path_factory.mktemp("combiner") / "combiner_stress")
    src = os.path.join(ROOT, "tests", "cpp", "combiner_stress.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", out, src], check=True)
    return out


@pytest.mark.parametrize("spin_us,threads,per_thread", [("400", 8, 600), ("0", 8, 600), ("400", 32, 150)])
def test_combiner_stress(stress_bin, spin_us, threads, per_thread):
    env = dict(os.environ, VL_COMBINE_SPIN_US=spin_us)
    r = subprocess.run([stress_bin, str(threads), str(per_thread)], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    rec = json.loads(r.stdout.strip().splitlines()[-1])
    assert rec["failures"] == 0 and rec["queries"] == threads * per_thread
    assert rec["combined"] > 0 and rec["max_batch"] > 1          # callers really were combined
    assert rec["max_batch"] <= threads // 2                       # two (metric) classes never share a batch
