# Round-2 GPU batch #24: conv2d_tc with elect-issued MMAs: backbone tests + timing
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/c24_*
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "backbone or channels_last or images_to_voxels" > gpurun_out/c24_tests.full 2>&1
tail -30 gpurun_out/c24_tests.full > gpurun_out/c24_tests.log; rm gpurun_out/c24_tests.full; tail -4 gpurun_out/c24_tests.log
timeout 300 python tools/dev/backbone_one.py 5 > gpurun_out/c24_backbone_time.log 2>&1; tail -2 gpurun_out/c24_backbone_time.log
