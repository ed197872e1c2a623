# Round-2 first GPU contact: full gpu suite (no -x), experimental backbone tests, default bench line.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.json
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/c1_gputests.log
OCC_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_backbone_gpu.py -q -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/c1_backbone.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
tail -25 gpurun_out/c1_gputests.log
tail -8 gpurun_out/c1_backbone.log
tail -c 1500 gpurun_out/c1_bench.json; tail -5 gpurun_out/c1_bench.err
