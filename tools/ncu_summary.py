"""Summarise an .ncu-rep (raw + source pages) into text: python tools/ncu_summary.py file.ncu-rep [kernel-index]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'smsp__warps_active.avg.per_cycle_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg',
        'sm__sass_thread_inst_executed_op_dfma_pred_on.sum', 'sm__sass_thread_inst_executed_op_dmul_pred_on.sum',
        'sm__sass_thread_inst_executed_op_dadd_pred_on.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('=' * 100)
    for k in KEYS:
        if k in d:
            print(f"{k:70s} {d[k]:>18s} {units[hdr.index(k)]}")
    st = {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''): float(v)
          for k, v in d.items() if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('ratio')}
    print("stalls (warps per issue):", ", ".join(f"{k}={v:.2f}" for k, v in sorted(st.items(), key=lambda x: -x[1])[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
his = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
for n, hi in enumerate(his):
    hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
    end = his[n + 1] - 1 if n + 1 < len(his) else len(rows)
    agg = {}; samples = 0
    for r in rows[hi + 1:end]:
        if len(r) < len(hdr):
            continue
        srcl = r[ix['Source']]
        toks = srcl.split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]
        def col(name):          # a kernel without shared-memory traffic has no wavefront columns on its source page
            try:
                return int(float(r[ix[name]] or 0)) if name in ix else 0
            except ValueError:
                return 0
        ns = col('# Samples'); samples += ns
        ie = col('Instructions Executed')
        wf = col('L1 Wavefronts Shared'); wi = col('L1 Wavefronts Shared Ideal')
        parts = op.split('.')
        key = parts[0] + ('.' + '.'.join(parts[1:3]) if parts[0] in ('LDS', 'STS', 'LDG', 'STG', 'LDGSTS') else '')
        a = agg.setdefault(key, [0, 0, 0, 0]); a[0] += ie; a[1] += ns; a[2] += wf; a[3] += wi
    print('-' * 100); print(rows[hi - 1][:2], "total samples", samples)
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:16]:
        print(f"  {k:18s} inst {v[0]:>10d} samples {v[1]:>7d} ({100*v[1]/max(1,samples):4.1f}%) smem wavefronts {v[2]:>10d} ideal {v[3]:>10d}")
