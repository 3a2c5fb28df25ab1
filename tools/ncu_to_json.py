"""Summarise an .ncu-rep (--set full) into a small JSON with units: python tools/ncu_to_json.py rep out.json [regex]"""
import csv, json, re, subprocess, sys
rep, outp = sys.argv[1], sys.argv[2]
pat = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active']
stalls = [h for h in hdr if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued')]
def to_bytes(v, u):
    m = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
    return float(v) * m.get(u, 1)
out = []
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    if pat and not pat.search(name): continue
    d = {"kernel": re.sub(r"\(.*", "", name)}
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            d[w] = f"{r[i]} {units[i]}".strip()
    if 'dram__bytes_read.sum' in hdr:
        i, j = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
        d["dram_bytes_per_launch"] = to_bytes(r[i], units[i]) + to_bytes(r[j], units[j])
    st = sorted(((float(r[hdr.index(h)] or 0), h.replace('smsp__pcsamp_warps_issue_stalled_', '')) for h in stalls), reverse=True)
    tot = sum(v for v, _ in st) or 1
    d["stall_share_pct"] = {h: round(100 * v / tot, 1) for v, h in st[:8]}
    out.append(d)
json.dump({"source": rep, "launches": out}, open(outp, "w"), indent=1)
print(len(out), "launches ->", outp)
