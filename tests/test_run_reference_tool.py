"""tools/run_reference.py (the reference arm through the real Cherab + Raysect API) must stay importable and must say cleanly
when Cherab / Raysect are not installed — so the real-reference path cannot rot unnoticed."""
import importlib.util
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load():
    spec = importlib.util.spec_from_file_location("run_reference", os.path.join(ROOT, "tools", "run_reference.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_run_reference_imports_and_probes():
    mod = _load()
    why = mod.available()
    assert why is None or (isinstance(why, str) and why)            # None: both packages import; else a one-line reason
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_reference.py"), "--probe"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["available"] == (why is None)


def test_real_reference_world_builds_when_cherab_is_installed():
    mod = _load()
    if mod.available() is not None:
        pytest.skip("Cherab / Raysect are not installed here: %s" % mod.available())
    world, plasma = mod._world(32, 650.0, 660.0)
    assert len(list(plasma.models)) == 9
