"""The oracle against the reference's own known-answer tests:
test_sparse_trace.f90, test_comparator.f90, test_piecewise_linear_function.f90, test_source_bilat.f90,
test_orthodrome.f90, test_euler.f90 restated in oracle/kat_main.cpp."""
import re
import subprocess

import oracle_lib


def test_reference_known_answer_tests():
    r = subprocess.run([oracle_lib.KAT_PATH], capture_output=True, text=True, timeout=300)
    m = re.search(r"kat: (\d+) checks, (\d+) failures", r.stdout)
    assert m, r.stdout + r.stderr
    assert int(m.group(1)) >= 60
    assert int(m.group(2)) == 0 and r.returncode == 0, r.stdout
