cd /root/repo
rm -f gpurun_out/_sweep_ref.npy gpurun_out/sweep.jsonl gpurun_out/sweep.err
run() { tag=$1; shift; env "$@" timeout 120 python tests/_sweep_gather.py $tag >> gpurun_out/sweep.jsonl 2>> gpurun_out/sweep.err; }
run conv96_head8
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/sweep_pytest.log
cat gpurun_out/sweep.jsonl
tail -5 gpurun_out/sweep.err
cat gpurun_out/sweep_pytest.log
