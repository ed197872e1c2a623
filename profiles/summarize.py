"""Turn ncu reports (gpurun_out/*.ncu-rep, scratch) into the small tracked summaries under profiles/.

    python profiles/summarize.py gpurun_out/prof_r1_layers.ncu-rep gpurun_out/prof_r1_tail.ncu-rep

Writes profiles/r1_ncu_summary.csv (one row per captured launch) and profiles/r1_traffic.json (DRAM bytes per launch
and per frame by kernel category, read by bench.py for `roofline.traffic`)."""
import csv
import io
import json
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed']
UNIT = {'Mbyte': 1e6, 'Kbyte': 1e3, 'Gbyte': 1e9, 'byte': 1.0, 'us': 1.0, 'ms': 1e3, 'ns': 1e-3, 'msecond': 1e3, 'usecond': 1.0}
CATEGORY = [('sca_', 'sca_gather'), ('tsa_fused', 'tsa_gather'), ('gemm_tc', 'gemm'), ('conv3d_tc', 'conv3d'),
            ('head_tc', 'occ_head'), ('pack_level', 'pack')]
LAUNCHES_PER_FRAME = {'sca_gather': 6, 'tsa_gather': 6, 'conv3d': 2, 'occ_head': 1}


def rows_of(rep):
    """Per-launch rows of an .ncu-rep, or of a `ncu -i X.ncu-rep --page raw --csv` export made on the GPU box (*.csv)."""
    if rep.endswith('.csv'):
        out = open(rep).read()
    else:
        out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    hdr, units = r[0], r[1]
    hdr = [h.split('.TriageCompute.')[-1] if '.TriageCompute.' in h else h for h in hdr]
    for row in r[2:]:
        d = {'kernel': re.sub(r'\(.*', '', row[hdr.index('Kernel Name')]).replace('void occ::<unnamed>::', '')}
        for k in KEYS:
            if k in hdr:
                v = row[hdr.index(k)].replace(',', '')
                try:
                    d[k] = float(v) * UNIT.get(units[hdr.index(k)], 1.0)
                except ValueError:
                    d[k] = v
        yield d


def traffic_of(rows, source, whole_frame=False):
    """DRAM bytes per launch / per frame by kernel category from the per-launch rows.  whole_frame: the rows are exactly
    one frame (every launch captured once), so per-frame numbers are plain sums."""
    traffic = {}
    for pat, cat in CATEGORY:
        sel = [d for d in rows if pat in d['kernel']]
        if not sel:
            continue
        per = [float(d['dram__bytes_read.sum']) + float(d['dram__bytes_write.sum']) for d in sel]
        dur = [float(d['gpu__time_duration.sum']) for d in sel]
        traffic[cat] = {'dram_bytes_per_launch': sum(per) / len(per), 'captured_launches': len(sel),
                        'avg_duration_us': sum(dur) / len(dur), 'source': source}
        if whole_frame:
            traffic[cat].update(dram_bytes_per_frame=sum(per), launches_per_frame=len(sel), duration_us_per_frame=sum(dur))
            continue
        if cat in LAUNCHES_PER_FRAME:
            traffic[cat]['dram_bytes_per_frame'] = traffic[cat]['dram_bytes_per_launch'] * LAUNCHES_PER_FRAME[cat]
        if cat == 'gemm' and len(sel) >= 8:
            # the capture holds the hoisted value_proj GEMM (largest) + the 7 GEMMs of one encoder layer: scale the
            # layer to 6 layers instead of averaging unlike launches
            big = max(range(len(sel)), key=lambda i: per[i])
            layer = [per[i] for i in range(len(sel)) if i != big][:7]
            frame = per[big] + 6 * sum(layer)
            traffic[cat].update(dram_bytes_per_frame=frame, dram_bytes_per_launch=frame / 43,
                                note='value_proj GEMM + 6 x (7 GEMMs of the captured layer), / 43 launches')
    return traffic


def main(reps):
    if reps and reps[0] == '--from-csv':                        # rebuild r1_traffic.json from the committed per-launch CSV
        rows = list(csv.DictReader(open(os.path.join(HERE, 'r1_ncu_summary.csv'))))
        src = reps[1] if len(reps) > 1 else 'r1_ncu_summary.csv'
        traffic = traffic_of(rows, src)
        json.dump(traffic, open(os.path.join(HERE, 'r1_traffic.json'), 'w'), indent=1)
        for k, v in traffic.items():
            print(k, {a: (round(b / 1e6, 1) if 'bytes' in a else b) for a, b in v.items() if a not in ('source', 'note')})
        return
    tag, whole = 'r1', False
    if reps and reps[0] == '--frame':                           # --frame TAG file: one whole frame captured (57 launches)
        tag, whole, reps = reps[1], True, reps[2:]
    rows = [d for rep in reps for d in rows_of(rep)]
    with open(os.path.join(HERE, tag + '_ncu_summary.csv'), 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['kernel'] + KEYS)
        for d in rows:
            w.writerow([d['kernel']] + [d.get(k, '') for k in KEYS])
    traffic = traffic_of(rows, ', '.join(os.path.basename(r) for r in reps), whole)
    json.dump(traffic, open(os.path.join(HERE, tag + '_traffic.json'), 'w'), indent=1)
    for k, v in traffic.items():
        print(k, {a: (round(b / 1e6, 1) if 'bytes' in a else b) for a, b in v.items() if a not in ('source', 'note')})


if __name__ == '__main__':
    main(sys.argv[1:])
