# Round-2 GPU batch #5: SM-tiled persistent SCA gather: parity (bf16 groups), A/B against the linear mapping, ncu of the new kernel.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.json gpurun_out/c5_*
run() { name=$1; shift; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider "$@" > gpurun_out/c5_tests_$name.full 2>&1; tail -70 gpurun_out/c5_tests_$name.full > gpurun_out/c5_tests_$name.log; rm gpurun_out/c5_tests_$name.full; echo "== $name: $(tail -1 gpurun_out/c5_tests_$name.log)"; grep -E "^(FAILED|ERROR)|Error:|assert " gpurun_out/c5_tests_$name.log | head -12; }
run bf16    -k "bf16_simt or bf16_tensor_cores or bf16_feature or forward_host or pipelined or layer0_tsa or full_size_properties"
run full16  -k "full_size_six_layers_bf16 or full_size_one_layer"
timeout 900 python tools/dev/ab.py tiled= linear=OCC_SCA_TILED:0 > gpurun_out/c5_ab.log 2>&1
cat gpurun_out/c5_ab.log | cut -c1-400
cp gpurun_out/ab.json gpurun_out/c5_ab.json
AB_FRAMES=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:sca_tile -s 2 -c 1 -o gpurun_out/c5_sca_tile python tools/dev/ab_one.py > /dev/null 2>&1
ncu -i gpurun_out/c5_sca_tile.ncu-rep --page raw --csv > gpurun_out/c5_sca_tile_raw.csv 2>/dev/null
python - <<'P'
import csv
r=list(csv.reader(open('gpurun_out/c5_sca_tile_raw.csv')))
h,u,v=r[0],r[1],r[2]
for k in ('gpu__time_duration.sum','l1tex__t_sector_hit_rate.pct','l1tex__m_xbar2l1tex_read_bytes.sum','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','dram__bytes_read.sum','lts__t_sector_hit_rate.pct','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active'):
    if k in h: print(k, v[h.index(k)], u[h.index(k)])
P
ls -la gpurun_out | grep c5_
