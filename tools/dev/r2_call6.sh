# Round-2 GPU batch #6: SM-tiled SCA gather with 5 CTAs/SM (no spills) vs 6 CTAs/SM (spills) vs linear
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/c6_*
timeout 900 python tools/dev/ab.py tiled6= tiled5=OCC_SCA_TILED:5 linear=OCC_SCA_TILED:0 > gpurun_out/c6_ab.log 2>&1
cat gpurun_out/c6_ab.log | cut -c1-400
cp gpurun_out/ab.json gpurun_out/c6_ab.json
OCC_SCA_TILED=5 AB_FRAMES=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:sca_tile -s 2 -c 1 -o gpurun_out/c6_sca_tile5 python tools/dev/ab_one.py > /dev/null 2>&1
ncu -i gpurun_out/c6_sca_tile5.ncu-rep --page raw --csv > gpurun_out/c6_sca_tile5_raw.csv 2>/dev/null
python - <<'P'
import csv
r=list(csv.reader(open('gpurun_out/c6_sca_tile5_raw.csv')))
h,u,v=r[0],r[1],r[2]
for k in ('gpu__time_duration.sum','l1tex__t_sector_hit_rate.pct','l1tex__m_xbar2l1tex_read_bytes.sum','l1tex__m_l1tex2xbar_write_bytes.sum','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio'):
    if k in h: print(k, v[h.index(k)], u[h.index(k)])
P
