#!/usr/bin/env python
"""Table of the key metrics of an `ncu --set full` raw CSV export (ncu -i x.ncu-rep --page raw --csv), one row per launch.

    python tools/ncu_csv_table.py gpurun_out/r2e/ncu_raw_f16f8.csv > profiles/r2_conv_tc_ncu_full_f16f8.md
"""
import csv, sys
KEYS = [("gpu__time_duration.sum","time"),("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","tensor%"),
 ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active","hmma%"),
 ("dram__bytes_read.sum","DRAM rd"),("dram__bytes_write.sum","DRAM wr"),("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","DRAM%"),
 ("l1tex__m_xbar2l1tex_read_bytes.sum","L2->SM"),("lts__t_sector_hit_rate.pct","L2hit%"),("lts__throughput.avg.pct_of_peak_sustained_elapsed","LTS%"),
 ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed","L1%"),("sm__throughput.avg.pct_of_peak_sustained_elapsed","SM%"),("launch__grid_size","grid"),
 ("sm__cycles_active.avg","cycles"),("smsp__cycles_active.avg","smsp cyc")]
def load(path):
    rows=list(csv.reader(open(path)))
    # find header row
    for i,r in enumerate(rows):
        if "Kernel Name" in r: hdr=r; units=rows[i+1]; data=rows[i+2:]; break
    return hdr,units,data
for path in sys.argv[1:]:
    hdr,units,data=load(path)
    have=[(k,n) for k,n in KEYS if k in hdr]
    print(path)
    print("| # | kernel | "+" | ".join(n for _,n in have)+" |")
    for i,d in enumerate(data):
        if len(d)<len(hdr): continue
        name=d[hdr.index("Kernel Name")]
        import re
        m=re.search(r"conv_tc_kernel<([^>]*)>",name); name=m.group(1) if m else name[:30]
        vals=[]
        for k,_ in have:
            j=hdr.index(k); v=d[j]; u=units[j]
            try:
                f=float(v.replace(",",""))
                if u in("byte",) : v="%.2f GB"%(f/1e9)
                elif u=="Mbyte": v="%.2f GB"%(f/1e3)
                elif u=="Gbyte": v="%.2f GB"%f
                elif u=="Kbyte": v="%.2f MB"%(f/1e3)
                elif u in("us","usecond"): v="%.3f ms"%(f/1e3)
                elif u in("ms","msecond"): v="%.3f ms"%f
                elif u in("ns","nsecond"): v="%.3f ms"%(f/1e6)
                else: v="%.1f"%f if abs(f)<1e5 else "%.3g"%f
            except ValueError: pass
            vals.append(v)
        print("| %d | %s | "%(i,name)+" | ".join(vals)+" |")
