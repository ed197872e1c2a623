cd "$(dirname "$0")/../.."
rm -f gpurun_out/_sweep_ref.npy gpurun_out/sweep.jsonl gpurun_out/sweep.err
run() { tag=$1; shift; env "$@" timeout 120 python tools/dev/sweep_gather.py $tag >> gpurun_out/sweep.jsonl 2>> gpurun_out/sweep.err; }
run posfold
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/sweep_pytest.log
OCC_H2D_SPLIT=0 timeout 300 python bench.py --no-cpu --steps 20 --warmup 3 > gpurun_out/bench_split0.json 2> gpurun_out/bench_split0.err
OCC_H2D_SPLIT=1 timeout 300 python bench.py --no-cpu --steps 20 --warmup 3 > gpurun_out/bench_split1.json 2> gpurun_out/bench_split1.err
cat gpurun_out/sweep.jsonl
tail -5 gpurun_out/sweep.err
cat gpurun_out/sweep_pytest.log
