cd /root/repo
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
OCC_H2D_SPLIT=1 timeout 300 python bench.py --no-cpu --steps 20 --warmup 3 > gpurun_out/bench_split1.json 2>/dev/null
OCC_H2D_SPLIT=3 timeout 300 python bench.py --no-cpu --steps 20 --warmup 3 > gpurun_out/bench_split3.json 2>/dev/null
OCC_H2D_SPLIT=4 timeout 300 python bench.py --no-cpu --steps 20 --warmup 3 > gpurun_out/bench_split4.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:gemm_tc|sca_pipe|tsa_fused|pack_levels" --launch-skip 56 -c 20 -f -o gpurun_out/prof_r1_layers python tests/_sweep_gather.py ncuA > gpurun_out/ncuA.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:conv3d_tc|head_tc|bev_to_voxel|t32_convert" --launch-skip 5 -c 5 -f -o gpurun_out/prof_r1_tail python tests/_sweep_gather.py ncuB > gpurun_out/ncuB.log 2>&1
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1
tail -c 600 gpurun_out/bench_final.json; cat gpurun_out/final_pytest.log; tail -2 gpurun_out/final_smoke.log
