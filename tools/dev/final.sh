cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
OCC_H2D_SPLIT=4 timeout 300 python bench.py --no-cpu --steps 20 --warmup 3 > gpurun_out/bench_split4.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none -k "regex:gemm_tc|sca_pipe|tsa_fused|pack_levels" --launch-skip 56 -c 11 -f -o gpurun_out/prof_r1_layers python tools/dev/sweep_gather.py ncuA > gpurun_out/ncuA.log 2>&1
timeout 600 ncu --set full --clock-control none -k "regex:conv3d_tc|head_tc" --launch-skip 3 -c 3 -f -o gpurun_out/prof_r1_tail python tools/dev/sweep_gather.py ncuB > gpurun_out/ncuB.log 2>&1
python profiles/summarize.py gpurun_out/prof_r1_layers.ncu-rep gpurun_out/prof_r1_tail.ncu-rep > gpurun_out/summarize.log 2>&1
cp profiles/r1_ncu_summary.csv profiles/r1_traffic.json gpurun_out/
rm -f gpurun_out/_sweep_ref.npy
find gpurun_out -size +20M -delete
ls -la gpurun_out
du -sh gpurun_out
tail -c 300 gpurun_out/bench_final.json
