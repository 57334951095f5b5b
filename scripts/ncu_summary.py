"""Summarise an .ncu-rep (read on the CPU box): key metrics per captured launch + top source lines."""
import csv, subprocess, sys, io, json
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'sm__cycles_elapsed.avg', 'launch__waves_per_multiprocessor']
out = {}
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        out[k + (" [%s]" % units[i] if units[i] else "")] = [r[i] for r in rows[2:]]
print(json.dumps(out, indent=1))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", "0", "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
if hi:
    hi = hi[0]; h = rows[hi]
    iinst, isamp = h.index("Instructions Executed"), h.index("# Samples")
    stall = [i for i, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x]
    agg, st = {}, {}
    for r in rows[hi + 1:]:
        try:
            ln = int(r[0]); inst = int(r[iinst]); samp = int(r[isamp])
        except Exception:
            continue
        a = agg.setdefault(ln, [0, 0, r[1][:100]]); a[0] += inst; a[1] += samp
        for i in stall:
            try: st[h[i]] = st.get(h[i], 0) + int(r[i])
            except Exception: pass
    tot = sum(a[0] for a in agg.values()) or 1; ts = sum(a[1] for a in agg.values()) or 1
    print("total warp-instructions (launch 0):", tot, " samples:", ts)
    for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
        print("%5d %6.2f%% inst %6.2f%% samp  %s" % (ln, 100 * a[0] / tot, 100 * a[1] / ts, a[2]))
    ss = sum(st.values()) or 1
    print("stall reasons:", ", ".join("%s %.1f%%" % (k[6:], 100 * v / ss) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
