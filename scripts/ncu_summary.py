"""Summarise an .ncu-rep: per kernel launch key metrics + top stall reasons; optional opcode mix / region profile.
usage: python scripts/ncu_summary.py REP [--source KERNEL_REGEX:IDX]"""
import csv, collections, re, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__grid_size', 'launch__block_size', 'smsp__warps_active.avg.per_cycle_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio']
stall = [k for k in h if 'issue_stalled' in k and k.endswith('per_issue_active.ratio')]
for r in rows[2:]:
    print('----', r[h.index('Kernel Name')][:70])
    for k in keys:
        if k in h:
            print(f"  {k:80s} {r[h.index(k)]} {units[h.index(k)]}")
    st = sorted(((float(r[h.index(k)]), k) for k in stall), reverse=True)[:7]
    print("  stalls/issue: " + ", ".join(f"{k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}" for v, k in st))
if len(sys.argv) > 3 and sys.argv[2] == "--source":
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f"::regex:{sys.argv[3]}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = rows[1]; si = h.index('Source'); ei = h.index('Instructions Executed'); sm = h.index('# Samples')
    tot = 0; byop = collections.Counter(); static = collections.Counter(); samp = collections.Counter(); seq = []
    for r in rows[2:]:
        if len(r) <= ei or not r[ei].isdigit():
            continue
        m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[si].strip())
        op = m.group(2).split('.')[0] if m else '?'
        e = int(r[ei]); tot += e; byop[op] += e; static[op] += 1; samp[op] += int(r[sm]); seq.append((e, int(r[sm])))
    print("executed", tot, "static", len(seq))
    for op, c in byop.most_common(16):
        print(f"  {op:10s} exec {c / tot * 100:5.1f}%  static {static[op]:6d}  samples {samp[op]}")
    ts = sum(s for _, s in seq); step = max(500, len(seq) // 24)
    print("  region profile (static index: executed share, sample share)")
    for i in range(0, len(seq), step):
        e = sum(x[0] for x in seq[i:i + step]); s_ = sum(x[1] for x in seq[i:i + step])
        print(f"  {i:6d} exec {e / tot * 100:5.1f}% samples {s_ / ts * 100:5.1f}%")
