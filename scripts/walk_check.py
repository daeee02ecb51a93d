"""One short GPU run of the paired-end path-support tests without pytest/torch start-up: writes gpurun_out/walk_check.log.
Usage on the GPU box: python scripts/walk_check.py"""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
log = open(os.path.join(ROOT, "gpurun_out", "walk_check.log"), "w")


def say(*a):
    print(*a, file=log, flush=True)
    print(*a, flush=True)


from oracle import pyoracle  # noqa: E402
pyoracle.build()
import tests.test_walk_gpu as T  # noqa: E402
import pathlib  # noqa: E402
import tempfile  # noqa: E402

ok = True
jobs = [("support", T.test_pair_support_matches_oracle, c) for c in ("two_chromosomes", "noisy", "noisy_ragged")]
jobs += [("split", T.test_split_and_simplify_match_oracle, c) for c in (("two_chromosomes", 5), ("two_chromosomes", 10 ** 6), ("noisy", 1), ("noisy", 3))]
for name, fn, arg in jobs:
    t0 = time.time()
    try:
        fn(1, *arg) if isinstance(arg, tuple) else fn(1, arg)
        say("PASS", name, arg, "%.2fs" % (time.time() - t0))
    except Exception:
        ok = False
        say("FAIL", name, arg)
        say(traceback.format_exc())
for name, fn in (("script", lambda: T.test_graph_simplifier_script(1, pathlib.Path(tempfile.mkdtemp()))), ("limits", lambda: T.test_range_limits(1))):
    try:
        fn()
        say("PASS", name)
    except Exception:
        ok = False
        say("FAIL", name)
        say(traceback.format_exc())
say("ALL PASS" if ok else "SOME FAILED")
