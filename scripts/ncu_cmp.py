import csv, sys
def load(p):
    rows=list(csv.reader(open(p)))
    return {h:(v,u) for h,u,v in zip(rows[0],rows[1],rows[2])}
ds=[load(p) for p in sys.argv[1:]]
pats=['gpu__time_duration.sum','warps_active.avg.per_cycle','smsp__issue_active.avg.pct','fp64_cycles_active.avg.pct_of_peak_sustained_active','thread_inst_executed_per_inst','smsp__inst_executed.sum','_per_issue_active.ratio','sass__inst_executed_local','sm__icc_request_hit_rate','gcc__average_cache_request_hit_rate','gcc__cache_requests_type_instruction.sum.pct','gcc__xbar2gcc_sectors.sum.pct','l1tex__t_sector_hit_rate.pct','inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','launch__registers','sm__inst_executed_pipe_fp64.sum ','smsp__inst_executed_pipe_fp64']
for k in ds[0]:
    if any(s in k for s in pats):
        print(f"{k:92s} "+" ".join(f"{d.get(k,('',''))[0]:>16s}" for d in ds)+" "+ds[0][k][1])
