# Round-2 GPU batch #12: the driver's own commands -- full gpu suite in ONE pytest process (-x), smoke(), default bench line.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.json gpurun_out/c12_*
timeout 1700 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/c12_gputests.full 2>&1
tail -40 gpurun_out/c12_gputests.full > gpurun_out/c12_gputests.log; rm gpurun_out/c12_gputests.full; tail -5 gpurun_out/c12_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c12_smoke.log 2>&1; tail -3 gpurun_out/c12_smoke.log
timeout 900 python bench.py > gpurun_out/c12_bench.json 2> gpurun_out/c12_bench.err
tail -c 400 gpurun_out/c12_bench.json; tail -3 gpurun_out/c12_bench.err
