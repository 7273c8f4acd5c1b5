import csv, collections, sys
raw, src, ntiles = sys.argv[1], sys.argv[2], float(sys.argv[3])
rows=list(csv.reader(open(raw)))
hdr=rows[0]
for k in ['gpu__time_duration.sum','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__inst_executed.sum','launch__grid_size','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']:
    if k in hdr: print("%-80s %s"%(k, rows[2][hdr.index(k)]))
rows=list(csv.reader(open(src)))
hdr=rows[1]; iS=hdr.index('Source'); iE=hdr.index('Instructions Executed'); iSamp=hdr.index('# Samples')
stall_cols=[i for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot=0; byop=collections.Counter(); data=[]; agg={}
for idx,r in enumerate(rows[2:]):
    try: n=int(r[iE]); sm=int(r[iSamp] or 0)
    except: continue
    op=r[iS].split()[0] if not r[iS].startswith('@') else r[iS].split()[1]
    byop[op.split('.')[0]]+=n; tot+=n
    st=sorted([(int(r[i] or 0),hdr[i]) for i in stall_cols if (r[i] or '0').isdigit()],reverse=True)[:1]
    data.append((sm,n,idx,r[iS][:64],st))
    for i in stall_cols:
        if (r[i] or '0').isdigit(): agg[hdr[i]]=agg.get(hdr[i],0)+int(r[i])
print("total warp instr", tot, "per warp-unit", tot/ntiles)
print(" ".join("%s:%.0f"%(op,n/ntiles) for op,n in byop.most_common(26)))
tots=sum(d[0] for d in data)
print("samples",tots, sorted(agg.items(), key=lambda kv:-kv[1])[:8])
for sm,n,idx,s,st in sorted(data,reverse=True)[:14]:
    print("%5d (%4.1f%%) exec %8d  #%4d %-64s %s"%(sm,100*sm/tots,n,idx,s,st))
