# Round-2 GPU batch #3: grouped gpu suite (independent processes), bench, fp32 A/B, ncu launch list + one full-frame capture
# reduced to CSV on the box (gpurun_out/ must stay < 64 MiB).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.json gpurun_out/c3_* gpurun_out/*.ncu-rep
run() { name=$1; shift; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider "$@" > gpurun_out/c3_tests_$name.full 2>&1; tail -70 gpurun_out/c3_tests_$name.full > gpurun_out/c3_tests_$name.log; rm gpurun_out/c3_tests_$name.full; echo "== $name: $(tail -1 gpurun_out/c3_tests_$name.log)"; grep -E "^(FAILED|ERROR)|Error:|assert " gpurun_out/c3_tests_$name.log | head -12; }
run op      -k "ms_deform or pillar_projection or pack_levels or render_forward or ray_metric or missing_parameter"
run fp32    -k "engine_fp32 and not tensor_core"
run fp32tc  -k "tensor_core_split or tcgen05_gemm"
run bf16    -k "bf16_simt or bf16_tensor_cores or bf16_feature or forward_host or pipelined or layer0_tsa or full_size_properties"
run full32  -k "full_size_six_layers_fp32"
run full16  -k "full_size_six_layers_bf16"
run plugin  -k "plugin or temporal or rotation or detector_output or detector_temporal"
run backbone -k "backbone or channels_last or images_to_voxels"
cat gpurun_out/parity_report.json 2>/dev/null | head -120
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err
tail -c 300 gpurun_out/c3_bench.json; tail -3 gpurun_out/c3_bench.err
timeout 900 python tools/dev/ab.py base= headmajor=OCC_VALUE_HEADMAJOR:1 fp32tc=AB_PRECISION:fp32,AB_TC:1,AB_FRAMES:40 > gpurun_out/c3_ab.log 2>&1
cat gpurun_out/c3_ab.log | cut -c1-400
AB_FRAMES=3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 228 -c 120 --csv --log-file gpurun_out/c3_launches.csv \
    python tools/dev/ab_one.py > gpurun_out/c3_ncu_list.log 2>&1
AB_FRAMES=3 timeout 900 ncu --set full --clock-control none -s 114 -c 57 -o /tmp/c3_prof_frame python tools/dev/ab_one.py > gpurun_out/c3_ncu_full.log 2>&1
ncu -i /tmp/c3_prof_frame.ncu-rep --page raw --csv > gpurun_out/c3_frame_raw.csv 2>/dev/null
AB_FRAMES=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:sca_pipe -s 2 -c 1 -o gpurun_out/c3_sca python tools/dev/ab_one.py > /dev/null 2>&1
ls -la gpurun_out | grep c3_; du -sh gpurun_out
