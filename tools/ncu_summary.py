"""Summarise an .ncu-rep (read offline with `ncu -i`) into a small text file for profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/out.txt [TRAFFIC_KEY]
With TRAFFIC_KEY (e.g. 2048x2048) the mean dram__bytes_read.sum + dram__bytes_write.sum per profiled sweep launch is written
to profiles/traffic.json under that key, with the summary file as its source (bench.py reports it as roofline.traffic)."""
import csv, json, os, re, subprocess, sys
from collections import defaultdict

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores', 'smsp__cycles_active.avg']


def short(name):
    m = re.search(r'LineCfg<\(int\)(\d+), \(int\)(\d+), \(int\)(\d+), \(int\)(\d+), \(int\)(\d+)>, \(int\)(\d+), \(bool\)(\d), \(int\)(\d), \(int\)(\d), \(int\)(\d)', name)
    if not m:
        m2 = re.search(r'(sweep_kernel|pipe_col_conv_kernel)<bdof::LineCfg<\(int\)(\d+), \(int\)(\d+), \(int\)(\d+), \(int\)(\d+), \(int\)(\d+)>, \(int\)(\d+), \(int\)(\d+)(?:, \(bool\)(\d), \(bool\)(\d))?', name)
        if m2:
            k, n, t, r1, r2, r3, lpc, parts, col, adj = m2.groups()
            kind = k if col is None else 'sweep_kernel %s %s' % ('y (columns)' if col == '1' else 'x (rows)', 'adjoint' if adj == '1' else 'forward')
            return '%s N=%s T=%s radices=%sx%s lines/CTA=%s exchange parts=%s' % (kind, n, t, r1, r2, lpc, parts)
        return name.split('(')[0]
    n, t, r1, r2, r3, lpc, col, mode, pre, post = m.groups()
    kind = ('col' if col == '1' else 'row') + '_' + ['conv', 'fft', 'ifft', 'conv2d'][int(mode)] + ('_transmit' if pre == '1' else '') + ('_adjoint' if post == '1' else '')
    return 'line_kernel N=%s T=%s radices=%sx%sx%s lines/CTA=%s %s' % (n, t, r1, r2, r3, lpc, kind)


def to_bytes(val, unit):
    v = float(val.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)


def main(rep, out, traffic_key=None):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    lines = ['ncu summary of %s (ncu --set full --clock-control none; per-launch values, cold cache, serialised)' % rep, '']
    seen = set()
    per_kernel = {}
    for i, d in enumerate(data):
        name = short(d[hdr.index('Kernel Name')])
        if 'sweep_kernel' in name and 'dram__bytes_read.sum' in hdr:
            r, w = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
            per_kernel.setdefault(name, []).append(to_bytes(d[r], units[r]) + to_bytes(d[w], units[w]))
        if name in seen:
            continue
        seen.add(name)
        lines.append('== ' + name)
        for k in KEYS:
            if k in hdr:
                lines.append('   %-70s %s %s' % (k, d[hdr.index(k)], units[hdr.index(k)]))
        # stall reasons from the source page
        src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-id', '::regex:kernel:%d' % (i + 1)], capture_output=True, text=True).stdout
        srows = list(csv.reader(src.splitlines()))
        tot = defaultdict(float)
        for j, r in enumerate(srows):
            if r and r[0] == 'Address':
                h2 = r
                cols = [k for k, h in enumerate(h2) if h.startswith('stall_') and 'Not Issued' not in h]
                for rr in srows[j + 1:]:
                    if len(rr) == len(h2):
                        for k in cols:
                            try:
                                tot[h2[k]] += float(rr[k])
                            except ValueError:
                                pass
                break
        s = sum(tot.values())
        if s:
            lines.append('   warp stall samples: ' + ', '.join('%s %.1f%%' % (k[6:], 100 * v / s) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:7]))
        lines.append('')
    open(out, 'w').write('\n'.join(lines))
    print('wrote', out)
    if traffic_key and per_kernel:
        path = os.path.join(os.path.dirname(os.path.abspath(out)), 'traffic.json')
        try:
            tj = json.load(open(path))
        except Exception:
            tj = {}
        means = {k: sum(v) / len(v) for k, v in per_kernel.items()}
        tj[traffic_key] = {'bytes_per_launch': sum(means.values()) / len(means), 'per_kernel': means,
                           'source': 'ncu --set full, %s (dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the %d sweep '
                                     'kernel instantiations profiled)' % (os.path.relpath(out, os.path.dirname(os.path.dirname(os.path.abspath(out)))), len(means))}
        json.dump(tj, open(path, 'w'), indent=1, sort_keys=True)
        print('updated', path)


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
