# Round-2 GPU batch: the gpu suite in independent pytest processes (a device fault in one group must not poison the rest),
# experimental backbone tests, default bench line, kernel-variant A/B, multi-lane experiment, ncu launch list.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.json gpurun_out/c2_*.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c2_smi.txt 2>&1
run() { name=$1; shift; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider "$@" 2>&1 | tail -40 > gpurun_out/c2_tests_$name.log; echo "== $name: $(tail -1 gpurun_out/c2_tests_$name.log)"; }
run op      -k "ms_deform or pillar_projection or pack_levels or render_forward or ray_metric or missing_parameter"
run fp32    -k "engine_fp32 and not tensor_core"
run fp32tc  -k "tensor_core_split or tcgen05_gemm"
run bf16    -k "bf16_simt or bf16_tensor_cores or bf16_feature or forward_host or pipelined or layer0_tsa or full_size_properties"
run full32  -k "full_size_six_layers_fp32"
run full16  -k "full_size_six_layers_bf16"
run plugin  -k "plugin or temporal or rotation or detector"
OCC_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_backbone_gpu.py -q -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/c2_backbone.log
echo "== backbone: $(tail -1 gpurun_out/c2_backbone.log)"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err
tail -c 600 gpurun_out/c2_bench.json; tail -3 gpurun_out/c2_bench.err
timeout 1500 python tools/dev/ab.py base= headmajor=OCC_VALUE_HEADMAJOR:1 \
    nol0=OCC_NO_L0_FOLD:1 conv1lane=OCC_CONV_SINGLE_LANE:1 fp32tc=AB_PRECISION:fp32,AB_TC:1,AB_FRAMES:40 \
    fp32simt=AB_PRECISION:fp32,AB_TC:0,AB_FRAMES:20 > gpurun_out/c2_ab.log 2>&1
cat gpurun_out/c2_ab.log | cut -c1-400
timeout 600 python tools/dev/lanes_exp.py > gpurun_out/c2_lanes.log 2>&1; cat gpurun_out/c2_lanes.log | tail -4
AB_FRAMES=3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 228 -c 120 --csv --log-file gpurun_out/c2_launches.csv \
    python tools/dev/ab_one.py > gpurun_out/c2_ncu_list.log 2>&1
tail -2 gpurun_out/c2_ncu_list.log
AB_FRAMES=3 timeout 900 ncu --set full --clock-control none --import-source on -s 114 -c 57 -o gpurun_out/c2_prof_frame \
    python tools/dev/ab_one.py > gpurun_out/c2_ncu_full.log 2>&1
tail -2 gpurun_out/c2_ncu_full.log; ls -la gpurun_out/*.ncu-rep
du -sh gpurun_out
