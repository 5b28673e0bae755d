"""Oracle vs the unmodified reference at full model sizes -- only where /root/reference exists (the authoring container).
On the GPU box and anywhere else the committed goldens (tests/test_oracle_golden.py) carry the pin."""
import os
import subprocess
import sys

import pytest

REF = '/root/reference'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'mvector')), reason='reference tree not present')
def test_oracle_bit_level_against_reference_full_size():
    env = dict(os.environ, PYTHONPATH=REF + os.pathsep + ROOT, OMP_NUM_THREADS='8')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'oracle_vs_reference_check.py')], env=env, cwd='/tmp',
                       capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0 and 'ORACLE_VS_REFERENCE_OK' in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
