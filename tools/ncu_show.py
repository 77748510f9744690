import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if "Metric Name" in r][0]
h=rows[hi]; mi,vi,ui=h.index("Metric Name"),h.index("Metric Value"),h.index("Metric Unit")
want=("Duration","DRAM Throughput","Memory Throughput","Issue Slots Busy","Achieved Occupancy","Theoretical Occupancy","Registers Per Thread","L2 Hit Rate","Executed Ipc Active","Warp Cycles Per Issued Instruction","Avg. Active Threads Per Warp")
for r in rows[hi+1:]:
    if len(r)>vi and r[mi] in want: print("  %-40s %s %s"%(r[mi],r[vi],r[ui]))
rows=list(csv.reader(open(sys.argv[2])))
hi=[i for i,r in enumerate(rows) if "ID" in r and "Kernel Name" in r][0]
h,u,v=rows[hi],rows[hi+1],rows[hi+2]
st=sorted(((float(v[i].replace(",","")) if v[i] not in ("","n/a") else 0.0,h[i]) for i in range(len(h)) if "issue_stalled" in h[i] and "average" in h[i] and h[i].endswith(".ratio")),reverse=True)[:7]
for val,k in st: print("   %8.2f %s"%(val,k.replace("smsp__average_warps_issue_stalled_","").replace("_per_issue_active.ratio","")))
for k in ("dram__bytes_read.sum","dram__bytes_write.sum","l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum","lts__t_sectors_op_atom.sum","lts__t_sectors_srcunit_tex_op_read.sum"):
    if k in h: print("  ",k,v[h.index(k)],u[h.index(k)])
