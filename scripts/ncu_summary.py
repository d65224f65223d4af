"""Summarise an exported ncu report: python scripts/ncu_summary.py raw.csv src.csv [window]"""
import csv, sys
raw, src = sys.argv[1], sys.argv[2]
W = int(sys.argv[3]) if len(sys.argv) > 3 else 400
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__icc_request_hit_rate.pct', 'idc__request_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct', 'smsp__inst_executed.sum', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'launch__registers_per_thread', 'lts__t_sector_hit_rate.pct',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum', 'sm__warps_active.avg.per_cycle_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print(h, u, v)
    elif 'smsp__average_warps_issue_stalled' in h and 'per_issue_active' in h:
        try:
            if float(v.replace(',', '')) > 0.05:
                print(h.replace('smsp__average_warps_issue_stalled_', 'stall ').replace('_per_issue_active.ratio', ''), v)
        except ValueError:
            pass
rows = list(csv.reader(open(src)))
hdr, data = rows[1], rows[2:]
iS, iSrc = hdr.index('# Samples'), hdr.index('Source')
tot = sum(int(r[iS]) for r in data)
print('total samples', tot, 'instrs', len(data))
for w in range(0, len(data), W):
    seg = data[w:w + W]
    s = sum(int(r[iS]) for r in seg)
    ops = {}
    for r in seg:
        t = r[iSrc].split()
        op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
        ops[op] = ops.get(op, 0) + 1
    print(w, s, round(100 * s / tot, 1), sorted(ops.items(), key=lambda x: -x[1])[:4])
for i in sorted(range(len(data)), key=lambda i: -int(data[i][iS]))[:14]:
    print(i, data[i][iS], data[i][iSrc].strip()[:60], '| prev:', data[i - 1][iSrc].strip()[:50])
