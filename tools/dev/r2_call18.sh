# Round-2 GPU batch #18: shared-memory stem im2col (backbone tests + timing), conv3d_tc ncu capture with source for stall attribution
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/c18_*
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "backbone or channels_last or images_to_voxels" > gpurun_out/c18_tests.full 2>&1
tail -30 gpurun_out/c18_tests.full > gpurun_out/c18_tests.log; rm gpurun_out/c18_tests.full; tail -4 gpurun_out/c18_tests.log
timeout 300 python tools/dev/backbone_one.py 5 > gpurun_out/c18_backbone_time.log 2>&1; tail -2 gpurun_out/c18_backbone_time.log
AB_FRAMES=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3d_tc -s 4 -c 2 -o gpurun_out/c18_conv3d python tools/dev/ab_one.py > /dev/null 2>&1
ls -la gpurun_out/c18_conv3d.ncu-rep
