#!/usr/bin/env python3
"""Regenerate profiles/validated_kernels_<tag>.json from build/*.o -- run it on the objects a green GPU
suite has just used:  python tools/validated_kernels.py r02a "profiles/r02a_pytest.txt (N passed)" """
import hashlib, json, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, note = sys.argv[1], sys.argv[2]
man = {}
for f in sorted(os.listdir(os.path.join(root, 'build'))):
    if not f.endswith('.o'):
        continue
    out = subprocess.run(['cuobjdump', '-sass', os.path.join(root, 'build', f)], capture_output=True, text=True).stdout
    d, cur = {}, None
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1); d[cur] = []; continue
        if cur and re.match(r'\s*/\*[0-9a-f]{4}\*/', line):
            d[cur].append(re.sub(r'/\*[0-9a-f]+\*/', '', line).strip())
    if d:
        man[f[:-2]] = {k: hashlib.sha256('\n'.join(v).encode()).hexdigest()[:24] for k, v in d.items()}
json.dump({'validated_by': note, 'hash': 'first 24 hex digits of sha256 over the SASS instruction stream',
           'kernels': man}, open(os.path.join(root, 'profiles', f'validated_kernels_{tag}.json'), 'w'), indent=1,
          sort_keys=True)
print({k: len(v) for k, v in man.items()})
