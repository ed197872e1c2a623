# Round-2 GPU batch #20: TMA-store LayerNorm epilogue in gemm_tc.cu: parity groups (incl. temporal: y + pos copy), timing
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/c20_*
run() { name=$1; shift; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider "$@" > gpurun_out/c20_tests_$name.full 2>&1; tail -70 gpurun_out/c20_tests_$name.full > gpurun_out/c20_tests_$name.log; rm gpurun_out/c20_tests_$name.full; echo "== $name: $(tail -1 gpurun_out/c20_tests_$name.log)"; grep -E "^(FAILED|ERROR)|Error:|assert " gpurun_out/c20_tests_$name.log | head -12; }
run bf16    -k "bf16_simt or bf16_tensor_cores or bf16_feature or forward_host or pipelined or layer0_tsa or full_size_properties or chained or voxel_lift or tcgen05_gemm"
run full16  -k "full_size_six_layers_bf16 or full_size_one_layer"
run plugin  -k "temporal or rotation or detector_temporal"
timeout 900 python tools/dev/ab.py tma_ln= > gpurun_out/c20_ab.log 2>&1
cat gpurun_out/c20_ab.log | cut -c1-400
