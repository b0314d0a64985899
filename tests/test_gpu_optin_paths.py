"""The opt-in conv3d paths (in-kernel split-K finish, dynamic tile scheduler) read their switches once per process, so
they are exercised by re-running the convolution / tokenizer parity tests in a child process with the switches set."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize('switch', ['OG_SPLITK_FUSED', 'OG_IGEMM_DYNAMIC'])
def test_conv_parity_with_optin_switch(switch):
    env = dict(os.environ, **{switch: '1'})
    r = subprocess.run([sys.executable, '-m', 'pytest', 'tests/test_gpu_layers.py', 'tests/test_gpu_tokenizer.py', '-m', 'gpu',
                        '-q', '-x', '-p', 'no:cacheprovider'], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
