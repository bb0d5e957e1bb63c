"""Turns gpurun_out/*.ncu-rep / launch CSVs into the small text summaries committed under profiles/."""
import collections, csv, json, subprocess, sys

def launches(csv_path, out):
    rows = [r for r in csv.reader(open(csv_path)) if len(r) > 5]
    hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[ix['Metric Name']] != 'gpu__time_duration.sum': continue
        name = r[ix['Kernel Name']].split('(')[0]
        v = float(r[ix['Metric Value']].replace(',', '')); u = r[ix['Metric Unit']]
        v *= {'us': 1e-3, 'ns': 1e-6, 's': 1e3, 'ms': 1.0}.get(u, 1.0)
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, 'w') as f:
        f.write(f"# per-kernel device time (ncu gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)\n")
        f.write(f"# source: {csv_path}\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{t:10.3f} ms {n:5d} launches {t / n:9.4f} ms/launch {100 * t / tot:5.1f}%  {k}\n")

KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic',
        'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.sum.per_cycle_elapsed',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio']

def full(rep, out, note=""):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, 'w') as f:
        f.write(f"# ncu --set full --clock-control none, source {rep}\n# {note}\n")
        for vals in rows[2:]:
            d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
            f.write(f"\n## {d.get('Kernel Name', '?')}  grid {d.get('launch__grid_size')} block {d.get('launch__block_size')}\n")
            for k in KEYS:
                if k in d: f.write(f"{k:90s} {d[k]:>18s} {u[k]}\n")
            src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
            srows = list(csv.reader(src.splitlines()))
            h2 = srows[1]; ix = {h: i for i, h in enumerate(h2)}
            ops = collections.Counter(); tot = 0
            for r in srows[2:]:
                if len(r) < len(h2): continue
                t = r[ix['Source']].split(); n = int(r[ix['Instructions Executed']] or 0)
                op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
                ops[op] += n; tot += n
            f.write(f"\n### executed warp instructions by opcode (total {tot})\n")
            for op, n in ops.most_common(24): f.write(f"{op:10s} {n:14d} {100 * n / tot:5.1f}%\n")
            break

if __name__ == "__main__":
    if sys.argv[1] == "launches": launches(sys.argv[2], sys.argv[3])
    else: full(sys.argv[2], sys.argv[3], " ".join(sys.argv[4:]))
