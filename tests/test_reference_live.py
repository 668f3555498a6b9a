"""Live cross-check of the oracle against the UNMODIFIED reference, run only where the reference tree is
present (the authoring container; skipped on the GPU box).  The bulk version is
oracle/validate_against_reference.py -> oracle/VALIDATION.json."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

REF = os.environ.get("MAACO_REF_DIR", "/root/reference")


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "MAACO.py")), reason="reference tree not present")
def test_oracle_matches_live_reference_sample():
    # separate interpreter: the reference's module names (MAACO, pso, ...) must not leak into this process
    out = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "validate_against_reference.py"), "120", "--no-write"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "'mismatches': 0" in out.stdout and "'path_mismatches': 0, 'stats_mismatches': 0" in out.stdout, out.stdout[-2000:]
