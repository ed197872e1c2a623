"""A/B runs of kernel variants (each in its own process): python tools/dev/ab.py name=ENV1:v,ENV2:v ... -> gpurun_out/ab.json"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
res = {}
for spec in sys.argv[1:]:
    name, _, envs = spec.partition('=')
    env = dict(os.environ)
    for kv in filter(None, envs.split(',')):
        k, _, v = kv.partition(':')
        env[k] = v
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'dev', 'ab_one.py')], env=env, capture_output=True, text=True,
                       timeout=600)
    line = [ln for ln in r.stdout.splitlines() if ln.startswith('{')]
    res[name] = json.loads(line[-1]) if line else {'error': (r.stderr or r.stdout)[-600:]}
    print(name, json.dumps(res[name]), flush=True)
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'ab.json'), 'w'), indent=1)
