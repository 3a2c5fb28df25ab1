"""One line per profiled launch of an .ncu-rep (dev tool): python tools/ncu_brief.py rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
def g(r, n, d="-"):
    return r[hdr.index(n)] if n in hdr else d
stalls = [h for h in hdr if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued')]
for r in rows[2:]:
    st = sorted(((float(r[hdr.index(h)] or 0), h.replace('smsp__pcsamp_warps_issue_stalled_', '')) for h in stalls), reverse=True)
    tot = sum(v for v, _ in st) or 1
    print("%-42s t=%8.1f us grid=%s regs=%s dmma=%5.1f%% fp64=%5.1f%% issue=%5.1f%% warps=%5.1f%% lts=%5.1f%% l1=%5.1f%% dramR=%s%s dramW=%s%s hit=%5.1f%%" % (
        g(r, 'Kernel Name')[:42], float(g(r, 'gpu__time_duration.sum')) * {"us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(rows[1][hdr.index('gpu__time_duration.sum')], 1),
        g(r, 'launch__grid_size'), g(r, 'launch__registers_per_thread'),
        float(g(r, 'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active', 0) or 0),
        float(g(r, 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 0) or 0),
        float(g(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active', 0) or 0),
        float(g(r, 'sm__warps_active.avg.pct_of_peak_sustained_active', 0) or 0),
        float(g(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 0) or 0),
        float(g(r, 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 0) or 0),
        g(r, 'dram__bytes_read.sum'), rows[1][hdr.index('dram__bytes_read.sum')], g(r, 'dram__bytes_write.sum'), rows[1][hdr.index('dram__bytes_write.sum')],
        float(g(r, 'lts__t_sector_hit_rate.pct', 0) or 0)))
    print("      stalls: " + ", ".join("%s %.0f%%" % (h, 100 * v / tot) for v, h in st[:7]))
