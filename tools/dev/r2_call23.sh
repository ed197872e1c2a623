# Round-2 GPU batch #23: conv3d with elect-issued MMAs (operands in the uniform datapath): parity + A/B vs the single-lane issue loop
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/c23_*
run() { name=$1; shift; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider "$@" > gpurun_out/c23_tests_$name.full 2>&1; tail -70 gpurun_out/c23_tests_$name.full > gpurun_out/c23_tests_$name.log; rm gpurun_out/c23_tests_$name.full; echo "== $name: $(tail -1 gpurun_out/c23_tests_$name.log)"; grep -E "^(FAILED|ERROR)|Error:|assert " gpurun_out/c23_tests_$name.log | head -12; }
run bf16    -k "bf16_simt or bf16_tensor_cores or full_size_properties or voxel_lift or pipelined"
run full16  -k "full_size_six_layers_bf16 or full_size_one_layer"
timeout 900 python tools/dev/ab.py head_elect= > gpurun_out/c23_ab.log 2>&1
cat gpurun_out/c23_ab.log | cut -c1-400
cp gpurun_out/ab.json gpurun_out/c23_ab.json
