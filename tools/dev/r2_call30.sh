# Round-2 GPU batch #30 (final-build evidence): full gpu suite in one process, smoke, bench, ncu launch list + whole-frame ncu --set full
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.json gpurun_out/c30_*
timeout 1700 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/c30_gputests.full 2>&1
tail -40 gpurun_out/c30_gputests.full > gpurun_out/c30_gputests.log; rm gpurun_out/c30_gputests.full; tail -5 gpurun_out/c30_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c30_smoke.log 2>&1; tail -3 gpurun_out/c30_smoke.log
timeout 900 python bench.py > gpurun_out/c30_bench.json 2> gpurun_out/c30_bench.err
tail -c 300 gpurun_out/c30_bench.json; tail -3 gpurun_out/c30_bench.err
AB_FRAMES=3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 208 -c 110 --csv --log-file gpurun_out/c30_launches.csv \
    python tools/dev/ab_one.py > gpurun_out/c30_ncu_list.log 2>&1
AB_FRAMES=3 timeout 900 ncu --set full --clock-control none -s 104 -c 52 -o /tmp/c30_prof_frame python tools/dev/ab_one.py > gpurun_out/c30_ncu_full.log 2>&1
ncu -i /tmp/c30_prof_frame.ncu-rep --page raw --csv > gpurun_out/c30_frame_raw.csv 2>/dev/null
ls -la gpurun_out | grep c30_
