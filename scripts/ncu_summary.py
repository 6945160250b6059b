"""Summarise an .ncu-rep: per kernel launch key metrics + top stall reasons; optional opcode mix with shared wavefronts per
opcode and, with a cubin, attribution of executed instructions / samples / shared wavefronts to SOURCE LINES (nvdisasm line
table x ncu SASS page, matched by instruction order).
usage: python scripts/ncu_summary.py REP [--source KERNEL_REGEX [--per N] [--cubin OBJ_OR_CUBIN --mangled SUBSTR]]"""
import argparse, collections, csv, io, os, re, subprocess, sys, tempfile

ap = argparse.ArgumentParser()
ap.add_argument("rep")
ap.add_argument("--source", default=None, help="kernel-name regex for the SASS page")
ap.add_argument("--per", type=float, default=1.0, help="divide counts by this (e.g. subproblems per launch)")
ap.add_argument("--cubin", default=None, help=".o / .cubin with -lineinfo that holds the kernel")
ap.add_argument("--mangled", default=None, help="substring of the mangled kernel name inside the cubin")
ap.add_argument("--file", default="ip_kernel.cuh", help="source file whose lines are listed")
ap.add_argument("--top", type=int, default=30)
args = ap.parse_args()

raw = subprocess.run(["ncu", "-i", args.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__grid_size', 'launch__block_size', 'smsp__warps_active.avg.per_cycle_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio']
stall = [k for k in h if 'issue_stalled' in k and k.endswith('per_issue_active.ratio')]
for r in rows[2:]:
    print('----', r[h.index('Kernel Name')][:90])
    for k in keys:
        if k in h:
            print(f"  {k:80s} {r[h.index(k)]} {units[h.index(k)]}")
    st = sorted(((float(r[h.index(k)]), k) for k in stall), reverse=True)[:7]
    print("  stalls/issue: " + ", ".join(f"{k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}" for v, k in st))
if not args.source:
    sys.exit(0)

src = subprocess.run(["ncu", "-i", args.rep, "--page", "source", "--csv", "--kernel-name", f"regex:{args.source}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r)
h = rows[hi]
si, ei, sm, wi = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples'), h.index('L1 Wavefronts Shared')
data = []
for r in rows[hi + 1:]:
    if len(r) > ei and r[ei].isdigit():
        data.append(r)
    elif data:
        break  # the page repeats per launch: keep the first
tot = sum(int(r[ei]) for r in data); ts = sum(int(r[sm]) for r in data); tw = sum(int(r[wi]) for r in data if r[wi].isdigit())
byop = collections.defaultdict(lambda: [0, 0, 0, 0])
for r in data:
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[si].strip())
    full = m.group(2) if m else '?'
    op = full.split('.')[0]
    if op in ('LDS', 'STS'):
        op = op + ('.128' if '128' in full else '.64' if '64' in full else '.32')
    b = byop[op]; b[0] += int(r[ei]); b[1] += 1; b[2] += int(r[sm]); b[3] += int(r[wi]) if r[wi].isdigit() else 0
print(f"\nSASS of {args.source}: static {len(data)} instructions, executed {tot / args.per:.1f}, shared wavefronts {tw / args.per:.1f} (per {args.per:g})")
for op, b in sorted(byop.items(), key=lambda kv: -kv[1][0])[:22]:
    print(f"  {op:10s} exec {b[0] / tot * 100:5.1f}% ({b[0] / args.per:8.1f})  static {b[1]:5d}  samples {b[2] / max(ts, 1) * 100:5.1f}%  wavefronts {b[3] / args.per:8.1f}")

if args.cubin and args.mangled:
    cub = args.cubin
    tmp = None
    if not cub.endswith(".cubin"):
        tmp = tempfile.mkdtemp()
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(cub)], cwd=tmp, capture_output=True)
        cub = os.path.join(tmp, [f for f in os.listdir(tmp) if f.endswith(".cubin")][0])
    dis = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout.split("\n")
    start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and args.mangled in l)
    end = next((i for i in range(start + 1, len(dis)) if dis[i].startswith("//--------------------- .text.")), len(dis))
    seq, cur = [], None
    for ln in dis[start:end]:
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
        if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', ln):
            seq.append(cur)
    if len(seq) != len(data):
        print(f"  (line table has {len(seq)} instructions, the SASS page {len(data)}: attribution skipped)")
        sys.exit(0)
    byline = collections.defaultdict(lambda: [0, 0, 0])
    for loc, r in zip(seq, data):
        b = byline[loc]; b[0] += int(r[ei]); b[1] += int(r[sm]); b[2] += int(r[wi]) if r[wi].isdigit() else 0
    print(f"\nsource lines by samples (top {args.top}):")
    for loc, b in sorted(byline.items(), key=lambda kv: -kv[1][1])[:args.top]:
        print(f"  {str(loc):34s} exec {b[0] / tot * 100:5.1f}%  samples {b[1] / max(ts, 1) * 100:5.1f}%  wavefronts {b[2] / max(tw, 1) * 100:5.1f}%")
    # by function of the listed file (device functions are recognised by their definition lines in the source)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "contactimplicitmpc.jl_b200", "csrc", args.file)
    if os.path.exists(path):
        defs = []
        for n, l in enumerate(open(path), 1):
            m = re.match(r'(?:__device__ __forceinline__|__global__)\s+[\w:<>,\s\*&]*?\b(\w+)\s*\(', l)
            if m:
                defs.append((n, m.group(1)))
        agg = collections.defaultdict(lambda: [0, 0, 0])
        for loc, b in byline.items():
            name = "(other files)" if not loc or loc[0] != args.file else next((f for n, f in reversed(defs) if n <= loc[1]), "(top)")
            if loc and loc[0] != args.file:
                name = loc[0]
            for i in range(3):
                agg[name][i] += b[i]
        print(f"\nby function of {args.file}:")
        for n, b in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            print(f"  {n:34s} exec {b[0] / tot * 100:5.1f}%  samples {b[1] / max(ts, 1) * 100:5.1f}%  wavefronts {b[2] / max(tw, 1) * 100:5.1f}%")
