"""Summarise ncu outputs brought back in gpurun_out/ into small text files for profiles/."""
import csv, collections, subprocess, sys, io

def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = None; agg = collections.OrderedDict()
    for r in rows:
        if r[0] == 'ID': hdr = r; continue
        if hdr is None: continue
        d = dict(zip(hdr, r))
        name = d['Kernel Name'].split('(')[0][-70:]
        agg.setdefault(name, []).append(float(d['Metric Value'].replace(',', '')))
    tot = sum(sum(v) for v in agg.values())
    out = [f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)", f"# total {tot/1e3:.1f} us over {sum(len(v) for v in agg.values())} launches"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"{k:70s} n={len(v):3d} mean={sum(v)/len(v)/1e3:9.1f} us share={sum(v)/tot*100:5.1f}%")
    return "\n".join(out)

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_static', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu.sum']

def full(rep):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    out = [f"# ncu --set full --clock-control none : {rep}"]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        out.append("== " + d['Kernel Name'].split('(')[0] + f"  grid={d.get('Grid Size','?')} block={d.get('Block Size','?')}")
        for w in WANT:
            if w in d: out.append(f"   {w:70s} {d[w]:>16s} {units[hdr.index(w)]}")
    return "\n".join(out)

if __name__ == "__main__":
    kind, src = sys.argv[1], sys.argv[2]
    print(launches(src) if kind == "launches" else full(src))
