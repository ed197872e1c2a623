# Round-2 GPU batch #4: bf16 parity tests with the recalibrated bars + one-layer tight tests, descriptor-swizzle effect,
# pair-fetch alignment experiments (OCC_PAIR_HACK: timing only), multi-lane experiment.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.json gpurun_out/c4_*
run() { name=$1; shift; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider "$@" > gpurun_out/c4_tests_$name.full 2>&1; tail -70 gpurun_out/c4_tests_$name.full > gpurun_out/c4_tests_$name.log; rm gpurun_out/c4_tests_$name.full; echo "== $name: $(tail -1 gpurun_out/c4_tests_$name.log)"; grep -E "^(FAILED|ERROR)|Error:|assert " gpurun_out/c4_tests_$name.log | head -12; }
run bf16    -k "bf16_simt or bf16_tensor_cores or bf16_feature or forward_host or pipelined or layer0_tsa or full_size_properties"
run full16  -k "full_size_six_layers_bf16 or full_size_one_layer"
run plugin  -k "plugin or temporal or rotation or detector_output or detector_temporal"
cat gpurun_out/parity_report.json 2>/dev/null | grep -E "full1|_vs_bf16_model|class_agree" | head -60
timeout 1200 python tools/dev/ab.py base= headmajor=OCC_VALUE_HEADMAJOR:1 hm_even=OCC_VALUE_HEADMAJOR:1,OCC_PAIR_HACK:1 \
    hm_dup=OCC_VALUE_HEADMAJOR:1,OCC_PAIR_HACK:2 > gpurun_out/c4_ab.log 2>&1
cat gpurun_out/c4_ab.log | cut -c1-400
cp gpurun_out/ab.json gpurun_out/c4_ab.json
timeout 600 python tools/dev/lanes_exp.py > gpurun_out/c4_lanes.log 2>&1; tail -4 gpurun_out/c4_lanes.log
