"""Summarise an ncu report (ncu --set full ... -o X) as JSON: per profiled launch the duration, DRAM bytes, pipe utilisation,
registers, grid, bank conflicts, L2 hit rate and the top stall reasons of the source page.

    python tools/ncu_summary.py gpurun_out/X.ncu-rep profiles/X_summary.json ["note"]
"""
import collections, csv, io, json, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else None
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv', '--kernel-name-base', 'demangled'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = {'gpu__time_duration.sum': 'duration', 'dram__bytes_read.sum': 'dram_bytes_read', 'dram__bytes_write.sum': 'dram_bytes_write',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active': 'dmma_pipe_pct_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active': 'fp64_pipe_pct_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue_slots_pct_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps_active_pct', 'launch__registers_per_thread': 'registers_per_thread',
        'launch__grid_size': 'grid', 'launch__block_size': 'block', 'lts__t_sector_hit_rate.pct': 'l2_hit_rate_pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum': 'smem_bank_conflicts', 'launch__occupancy_limit_registers': 'occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem': 'occupancy_limit_shared_mem', 'smsp__inst_executed.sum': 'warp_instructions'}
mult = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'msecond': 1e-3, 'usecond': 1e-6, 'nsecond': 1e-9, 'second': 1.0,
        'ms': 1e-3, 'us': 1e-6, 'ns': 1e-9, 's': 1.0}
launches = []
for r in rows[2:]:
    d = {'kernel': r[hdr.index('Kernel Name')]}
    for k, name in want.items():
        if k in hdr:
            i = hdr.index(k)
            try:
                v = float(r[i].replace(',', ''))
            except ValueError:
                continue
            u = units[i]
            if u in mult and ('bytes' in k or 'duration' in k):
                v *= mult[u]
                name_u = name + ('_s' if 'duration' in k else '')
            else:
                name_u = name
            d[name_u] = v
    if 'dram_bytes_read' in d:
        d['dram_bytes_total'] = d['dram_bytes_read'] + d.get('dram_bytes_write', 0.0)
    launches.append(d)
# stall reasons from the source page
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name-base', 'demangled'], capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1] if len(r) > 1 else '', 'rows': []}; blocks.append(cur); continue
    if cur is not None:
        cur['rows'].append(r)
# (the CSV may list every launch twice: SASS view and source view) -> keep the first block with stall columns per launch, in order
usable = []
for b in blocks:
    if not b['rows']:
        continue
    h = b['rows'][0]
    if '# Samples' in h and 'Source' in h and 'Instructions Executed' in h:
        data = [r for r in b['rows'][1:] if len(r) >= len(h)]
        if data and (not usable or usable[-1][0] != b['name'] or len(usable) < len(launches)):
            usable.append((b['name'], h, data))
dedup = []
for u in usable:
    if dedup and dedup[-1][0] == u[0] and len(dedup[-1][2]) == len(u[2]):
        continue
    dedup.append(u)
for li, (name, h, data) in enumerate(dedup[:len(launches)]):
    cols = [i for i, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x]
    tot = collections.Counter()
    for r in data:
        for i in cols:
            tot[h[i]] += int(r[i] or 0)
    total = sum(tot.values())
    launches[li]['stall_samples_top'] = {k: round(v / total, 3) for k, v in tot.most_common(6)} if total else {}
    ops = collections.Counter()
    si, ie = h.index('Source'), h.index('Instructions Executed')
    for r in data:
        tok = r[si].split()
        op = (tok[1] if tok and tok[0].startswith('@') and len(tok) > 1 else (tok[0] if tok else '?')).split('.')[0]
        ops[op] += int(r[ie] or 0)
    launches[li]['executed_warp_instructions_top'] = dict(ops.most_common(8))
res = {'report': rep.split('/')[-1], 'tool': 'ncu --set full --clock-control none --import-source on (one launch each, cold cache, serialised)',
       'launches': launches}
if note:
    res['note'] = note
json.dump(res, open(out, 'w'), indent=1)
print(out)
