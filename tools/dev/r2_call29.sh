# Round-2 GPU batch #29: fp32 storage + tensor cores: Conv3d as three bf16-split passes (launch_conv3d_tc_split): parity + timing
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/c29_*
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "tensor_core_split or full_size_six_layers_fp32 or engine_fp32" > gpurun_out/c29_tests.full 2>&1
tail -40 gpurun_out/c29_tests.full > gpurun_out/c29_tests.log; rm gpurun_out/c29_tests.full; tail -8 gpurun_out/c29_tests.log
timeout 600 python tools/dev/ab.py fp32tc_conv_split=AB_PRECISION:fp32,AB_TC:1,AB_FRAMES:40 fp32tc_conv_simt=AB_PRECISION:fp32,AB_TC:1,AB_FRAMES:40,OCC_CONV_F32_SIMT:1 > gpurun_out/c29_ab.log 2>&1
cat gpurun_out/c29_ab.log | cut -c1-400
cp gpurun_out/ab.json gpurun_out/c29_ab.json
