# One gpurun batch for the first GPU contact of the backbone: parity tests of the explicit-im2col
# version, then of the implicit-GEMM kernel, then the experimental bench leg.  Keeps gpurun_out/ small.
#   gpurun --timeout 900 -- 'bash tools/dev/backbone_check.sh'
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
OCC_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_backbone_gpu.py -x -q 2>&1 | tail -30 > gpurun_out/backbone_tests.log
OCC_EXPERIMENTAL=1 OCC_BACKBONE_IMPLICIT=1 timeout 600 python -m pytest tests/test_backbone_gpu.py -x -q -k tcgen05 2>&1 | tail -30 > gpurun_out/backbone_tests_implicit.log
timeout 600 python bench.py --no-cpu --steps 10 --warmup 3 --with-backbone > gpurun_out/bench_backbone.json 2> gpurun_out/bench_backbone.err
OCC_BACKBONE_IMPLICIT=1 timeout 600 python bench.py --no-cpu --steps 10 --warmup 3 --with-backbone > gpurun_out/bench_backbone_implicit.json 2> gpurun_out/bench_backbone_implicit.err
tail -5 gpurun_out/backbone_tests.log gpurun_out/backbone_tests_implicit.log
tail -c 400 gpurun_out/bench_backbone.json; tail -c 400 gpurun_out/bench_backbone_implicit.json
du -sh gpurun_out
