import csv, sys
def f(x):
    try: return float(x.replace(',',''))
    except: return 0.0
KEYS=["gpu__time_duration.sum","launch__grid_size","launch__block_size","launch__registers_per_thread","sm__throughput.avg.pct_of_peak_sustained_elapsed","dram__throughput.avg.pct_of_peak_sustained_elapsed","dram__bytes_read.sum","dram__bytes_write.sum","lts__t_bytes.sum","sm__warps_active.avg.pct_of_peak_sustained_active","sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active","smsp__issue_active.avg.pct_of_peak_sustained_active","l1tex__throughput.avg.pct_of_peak_sustained_elapsed","lts__throughput.avg.pct_of_peak_sustained_elapsed","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","smsp__inst_executed.sum","sm__cycles_elapsed.max","smsp__inst_executed_pipe_xu.sum","sm__inst_executed_pipe_xu.sum"]
for fn in sys.argv[1:]:
    base=fn.replace('.ncu-rep','')
    rows=list(csv.reader(open(base+'.raw.csv')))
    hdr=rows[0]; units=rows[1]
    for r in rows[2:]:
        print("=====",base, r[hdr.index("Kernel Name")][:40])
        for k in KEYS:
            if k in hdr:
                i=hdr.index(k); print(f"  {k} = {r[i]} {units[i]}")
    try: rows=list(csv.reader(open(base+'.source.csv')))
    except Exception: continue
    hdr=rows[1]
    si=hdr.index("# Samples"); src=hdr.index("Source")
    stalls=[i for i,h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data=[r for r in rows[2:] if len(r)==len(hdr)]
    T=sum(f(r[si]) for r in data)
    print("  total samples",T)
    print("  "+"  ".join(f"{hdr[i][6:]}={100*sum(f(r[i]) for r in data)/T:.1f}%" for i in stalls if sum(f(r[i]) for r in data)/T>0.01))
    for r in sorted(data,key=lambda r:-f(r[si]))[:14]:
        st=sorted([(f(r[i]),hdr[i][6:]) for i in stalls],reverse=True)[:2]
        print(f"  {f(r[si]):7.0f} {r[src][:70]:70s} {st}")
