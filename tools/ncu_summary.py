import csv, collections, sys
raw, src = sys.argv[1], sys.argv[2]
rows=list(csv.reader(open(raw)))
hdr=rows[0]; vals=rows[2]
for k in ['gpu__time_duration.sum','smsp__inst_executed.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','smsp__cycles_active.avg','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active']:
    if k in hdr: print(k, vals[hdr.index(k)], rows[1][hdr.index(k)])
rows=list(csv.reader(open(src)))
hi=[i for i,r in enumerate(rows) if r and r[0]=="Address"][0]
hdr=rows[hi]
stall_cols=[i for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
samp=hdr.index('# Samples'); ie=hdr.index('Instructions Executed')
tot=collections.Counter(); total=0; data=[]; seen=set()
for r in rows[hi+1:]:
    if len(r)<len(hdr) or r[0] in seen: continue
    seen.add(r[0])
    try: n=int(r[samp])
    except: continue
    total+=n
    top=max(stall_cols, key=lambda i:int(r[i] or 0))
    for i in stall_cols:
        try: tot[hdr[i]]+=int(r[i])
        except: pass
    data.append((n,r[1][:60],r[ie],hdr[top]))
print("total samples", total, "distinct instr", len(data))
for k,v in tot.most_common(8): print(f"  {k:28s} {v:7d} {100*v/total:5.1f}%")
for n,s,e,t in sorted(data, reverse=True)[:int(sys.argv[3]) if len(sys.argv)>3 else 22]: print(f"  {n:6d} {100*n/total:5.1f}%  {s:60s} exec={e} {t}")
