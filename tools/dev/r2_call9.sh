# Round-2 GPU batch #9: merged TSA input GEMMs (multi-problem launch): parity, A/B, in-kernel timeline of one frame
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/c9_*
run() { name=$1; shift; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider "$@" > gpurun_out/c9_tests_$name.full 2>&1; tail -70 gpurun_out/c9_tests_$name.full > gpurun_out/c9_tests_$name.log; rm gpurun_out/c9_tests_$name.full; echo "== $name: $(tail -1 gpurun_out/c9_tests_$name.log)"; grep -E "^(FAILED|ERROR)|Error:|assert " gpurun_out/c9_tests_$name.log | head -12; }
run tcgemm  -k "tcgen05_gemm or tensor_core_split"
run bf16    -k "bf16_simt or bf16_tensor_cores or bf16_feature or forward_host or pipelined or layer0_tsa or full_size_properties"
run full16  -k "full_size_six_layers_bf16 or full_size_one_layer"
run plugin  -k "temporal or rotation or detector_temporal"
timeout 900 python tools/dev/ab.py merged= separate=OCC_TSA_MERGE:0 > gpurun_out/c9_ab.log 2>&1
cat gpurun_out/c9_ab.log | cut -c1-400
cp gpurun_out/ab.json gpurun_out/c9_ab.json
timeout 300 python tools/dev/gemm_timeline.py 2> gpurun_out/c9_timeline_merged.log > /dev/null

grep -c timeline gpurun_out/c9_timeline_merged.log gpurun_out/c9_timeline_separate.log
