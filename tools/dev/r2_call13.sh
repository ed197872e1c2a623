# Round-2 GPU batch #13: first GPU contact of the TMA-im2col implicit-GEMM conv2d (OCC_BACKBONE_IMPLICIT=1): backbone parity tests, then
# the images_to_voxels leg with and without it
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/c13_*
OCC_BACKBONE_IMPLICIT=1 timeout 600 python -m pytest tests/test_backbone_gpu.py -q -p no:cacheprovider > gpurun_out/c13_backbone_implicit.full 2>&1
tail -60 gpurun_out/c13_backbone_implicit.full > gpurun_out/c13_backbone_implicit.log; rm gpurun_out/c13_backbone_implicit.full; tail -25 gpurun_out/c13_backbone_implicit.log
python - <<'P' > gpurun_out/c13_backbone_bench.log 2>&1
import os, sys, json, subprocess
for flag in ('0', '1'):
    env = dict(os.environ); 
    if flag == '1': env['OCC_BACKBONE_IMPLICIT'] = '1'
    r = subprocess.run([sys.executable, 'bench.py', '--steps', '3', '--warmup', '3', '--repeats', '1', '--no-cpu', '--no-eager', '--no-dropin'], env=env, capture_output=True, text=True, timeout=900)
    line = [l for l in r.stdout.splitlines() if l.startswith('{')]
    if line:
        d = json.loads(line[-1]); print('implicit', flag, json.dumps(d.get('images_to_voxels')))
    else:
        print('implicit', flag, 'FAILED', r.stderr[-1500:])
P
cat gpurun_out/c13_backbone_bench.log | cut -c1-900
