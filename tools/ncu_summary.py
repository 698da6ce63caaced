#!/usr/bin/env python
"""Summaries of ncu outputs for profiles/: launch-list shares of one bench step, and key metrics of full captures."""
import csv, subprocess, sys, re, collections

def launch_list(path, steps_skip_last=True):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    iname, imetric, ival = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    iunit = hdr.index("Metric Unit")
    launches = []
    for r in rows[1:]:
        if r[imetric] != "gpu__time_duration.sum":
            continue
        v = float(r[ival].replace(",", ""))
        u = r[iunit]
        us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
        name = re.sub(r"\(.*", "", r[iname]).replace("void frb::", "").replace("frb::", "")
        launches.append((name, us))
    return launches

def last_step(launches):
    # one step = from a preprocess_u8_kernel launch to the next one; take the last complete step
    idx = [i for i, (n, _) in enumerate(launches) if n.startswith("preprocess_u8_kernel")]
    if len(idx) < 2:
        return launches
    a, b = idx[-2], idx[-1]
    return launches[a:b]

def raw(path, wanted):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    res = {"kernel": vals[hdr.index("Kernel Name")][:70]}
    for h, u, v in zip(hdr, units, vals):
        if h in wanted:
            res[h] = (v, u)
    return res

if __name__ == "__main__":
    if sys.argv[1] == "launches":
        step = last_step(launch_list(sys.argv[2]))
        tot = sum(us for _, us in step)
        agg = collections.OrderedDict()
        for n, us in step:
            a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += us
        print(f"one step: {len(step)} launches, {tot:.0f} us under ncu")
        print("| kernel | launches | us | share |\n|---|---|---|---|")
        for n, (c, us) in agg.items():
            print(f"| `{n}` | {c} | {us:.1f} | {100 * us / tot:.1f}% |")
    else:
        W = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
             "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed",
             "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
             "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_bytes.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
             "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
             "launch__grid_size", "launch__registers_per_thread", "gpc__cycles_elapsed.avg.per_second",
             "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
             "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__cycles_active.avg"]
        for pth in sys.argv[1:]:
            r = raw(pth, W)
            print("==", pth, r.pop("kernel"))
            for k, (v, u) in r.items():
                print(f"   {k}: {v} {u}")
