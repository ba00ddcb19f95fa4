"""Every kernel of the build that passed the GPU suite on a B200 (profiles/r01i_pytest.txt) is still in
the current build with a bit-identical SASS instruction stream.  Kernels may be ADDED (the features
written after the round's GPU minutes were spent add kernels and host code); a kernel of the validated
path may not change without a new hardware run -- in which case profiles/validated_kernels_*.json is
regenerated from the objects that run used (tools/validated_kernels.py)."""
import hashlib
import json
import os
import re
import shutil
import subprocess

import pytest


def _streams(obj):
    out = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True, check=True).stdout
    d, cur = {}, None
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1)
            d[cur] = []
            continue
        if cur and re.match(r'\s*/\*[0-9a-f]{4}\*/', line):
            d[cur].append(re.sub(r'/\*[0-9a-f]+\*/', '', line).strip())
    return {hashlib.sha256('\n'.join(v).encode()).hexdigest()[:24] for v in d.values()}


@pytest.mark.skipif(shutil.which('cuobjdump') is None, reason='needs the CUDA toolkit')
def test_validated_kernels_are_unchanged(repo_root):
    from hipace_b200.build import build_library
    build_library()
    man = json.load(open(os.path.join(repo_root, 'profiles', 'validated_kernels_r01i.json')))['kernels']
    n = 0
    for name, kernels in man.items():
        have = _streams(os.path.join(repo_root, 'build', name + '.o'))
        for kernel, h in kernels.items():
            assert h in have, f'{name}.cu: {kernel} differs from the hardware-validated build'
            n += 1
    assert n > 90
