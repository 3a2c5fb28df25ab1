"""Print selected raw metrics per kernel launch from an .ncu-rep (dev tool): python tools/ncu_rows.py rep [regex]"""
import csv, subprocess, sys, re
rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__cycles_active.avg', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active']
stalls = [h for h in hdr if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued')]
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    if pat and not pat.search(name): continue
    print("==", name[:110])
    for w in want:
        if w in hdr: print("   %-80s %s" % (w, r[hdr.index(w)]))
    st = sorted(((float(r[hdr.index(h)] or 0), h.replace('smsp__pcsamp_warps_issue_stalled_', '')) for h in stalls), reverse=True)
    tot = sum(v for v, _ in st) or 1
    print("   stalls: " + ", ".join("%s %.0f%%" % (h, 100 * v / tot) for v, h in st[:9]))
