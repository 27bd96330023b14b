"""CPU-only: the host-side concurrency of the search path, compiled with g++ against stand-in searches.
 * the caller combiner (csrc/combiner.h — concurrent single-query callers → one batched search): every caller gets
   the answer to its own query, failures stay with their caller, callers of different (k, metric, ef) classes never
   share a batch, nothing hangs — with polling waiters and without;
 * the shard group (csrc/group.cpp — fan-out to helper threads, completion hand-shake, stable merge in shard order)
   over 2 / 4 / 8 stand-in shards, polling and blocking;
 * both once more under ThreadSanitizer when the toolchain has it."""
import json
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
GROUP_SRC = [os.path.join(CPP, "group_stress.cpp"), os.path.join(ROOT, "vectorlite_b200", "csrc", "group.cpp")]


def _compile(out, sources, extra=()):
    return subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", *extra, "-o", out, *sources],
                          capture_output=True, text=True)


@pytest.fixture(scope="module")
def bins(tmp_path_factory):
    if not shutil.which("g++"):
        pytest.skip("needs g++")
    d = tmp_path_factory.mktemp("host_concurrency")
    out = {"combiner": str(d / "combiner_stress"), "group": str(d / "group_stress")}
    r = _compile(out["combiner"], [os.path.join(CPP, "combiner_stress.cpp")])
    assert r.returncode == 0, r.stderr
    r = _compile(out["group"], GROUP_SRC)
    assert r.returncode == 0, r.stderr
    tsan = {"combiner": str(d / "combiner_tsan"), "group": str(d / "group_tsan")}
    ok = (_compile(tsan["combiner"], [os.path.join(CPP, "combiner_stress.cpp")], ("-g", "-fsanitize=thread")).returncode == 0 and
          _compile(tsan["group"], GROUP_SRC, ("-g", "-fsanitize=thread")).returncode == 0)
    out["tsan"] = tsan if ok else None
    return out


def _run(binary, args, env_extra, timeout=180):
    env = dict(os.environ, **env_extra)
    r = subprocess.run([binary, *map(str, args)], env=env, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ThreadSanitizer" not in r.stderr, r.stderr
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("spin_us,threads,per_thread", [("400", 8, 600), ("0", 8, 600), ("400", 32, 150)])
def test_combiner_stress(bins, spin_us, threads, per_thread):
    rec = _run(bins["combiner"], [threads, per_thread], {"VL_COMBINE_SPIN_US": spin_us})
    assert rec["failures"] == 0 and rec["queries"] == threads * per_thread
    assert rec["combined"] > 0 and rec["max_batch"] > 1          # callers really were combined
    assert rec["max_batch"] <= threads // 2                       # two (metric) classes never share a batch


@pytest.mark.parametrize("shards", [2, 4, 8])
@pytest.mark.parametrize("spin_us", ["400", "0"])
def test_shard_group_stress(bins, shards, spin_us):
    rec = _run(bins["group"], [shards, 16, 250], {"VL_GROUP_SPIN_US": spin_us, "VL_COMBINE_SPIN_US": spin_us})
    assert rec["failures"] == 0 and rec["queries"] == 16 * 250
    # every combined batch reaches every shard exactly once: far fewer shard calls than queries x shards
    assert rec["shard_calls"] < rec["queries"] * shards


def test_host_concurrency_under_thread_sanitizer(bins):
    if not bins["tsan"]:
        pytest.skip("g++ without -fsanitize=thread")
    rec = _run(bins["tsan"]["combiner"], [8, 200], {"TSAN_OPTIONS": "halt_on_error=0"})
    assert rec["failures"] == 0
    for shards in (2, 8):
        rec = _run(bins["tsan"]["group"], [shards, 8, 80], {"TSAN_OPTIONS": "halt_on_error=0"})
        assert rec["failures"] == 0
