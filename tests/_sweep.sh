cd /root/repo
rm -f gpurun_out/_sweep_ref.npy gpurun_out/sweep.jsonl gpurun_out/sweep.err
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/sweep_pytest.log
run() { tag=$1; shift; env "$@" timeout 120 python tests/_sweep_gather.py $tag >> gpurun_out/sweep.jsonl 2>> gpurun_out/sweep.err; }
run new
run new_again
run sca813 OCC_SCA_PIPE=813
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -c 1 -f -o gpurun_out/prof_vgemm python tests/_sweep_gather.py ncu1 > gpurun_out/ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sca_pipe -c 1 -f -o gpurun_out/prof_sca_pipe python tests/_sweep_gather.py ncu2 > gpurun_out/ncu2.log 2>&1
cat gpurun_out/sweep.jsonl
tail -5 gpurun_out/sweep.err
cat gpurun_out/sweep_pytest.log
