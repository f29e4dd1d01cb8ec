"""Turn ncu captures brought back in gpurun_out/ into the small tracked summaries under profiles/.

  python profiles/summarize.py <tag> <launches.csv> <full.ncu-rep> <blocks-in-capture>
writes profiles/<tag>_launches.csv (the gpu__time_duration launch list), profiles/<tag>_stage_kernel.txt (key metrics,
stall breakdown and hottest SASS lines of every captured stage_kernel launch) and refreshes
profiles/stage_kernel_traffic.json (DRAM bytes per block per launch, read by bench.py for roofline.traffic).
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'lts__t_sector_hit_rate.pct', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic']


def main():
    tag, launches, rep, nblocks = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
    if launches != "-":
        shutil.copy(launches, os.path.join(HERE, f"{tag}_launches.csv"))
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = [f"# ncu --set full --clock-control none, {len(data)} captured launches of stage_kernel, {nblocks} blocks per launch", ""]
    names = [r[hdr.index('Kernel Name')] for r in data]
    out.append("kernel: " + names[0])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            out.append(f"{k} [{units[i]}]: " + ", ".join(r[i] for r in data))
    out.append("")
    out.append("warp stall reasons (warps stalled per issue-active cycle):")
    for i, k in enumerate(hdr):
        if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and 'not_issued' not in k:
            v = [float(r[i]) for r in data]
            if max(v) > 0.2:
                out.append("  " + k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '') + ": "
                           + ", ".join('%.2f' % x for x in v))
    ir, iw = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
    scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}
    tot = [float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]] for r in data]
    per_launch = sum(tot) / len(tot)
    out.append("")
    out.append(f"DRAM bytes per launch (read+write), mean over captured launches: {per_launch:.4e}  = {per_launch / nblocks:.1f} B per block")
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    ks = [i for i, r in enumerate(srows) if r and r[0] == 'Kernel Name']
    for n, k0 in enumerate(ks):
        seg = srows[k0:ks[n + 1]] if n + 1 < len(ks) else srows[k0:]
        h, body = seg[1], seg[2:]
        iS, iSrc, iE = h.index('# Samples'), h.index('Source'), h.index('Instructions Executed')
        st = [k for k in h if k.startswith('stall_') and 'Not Issued' not in k]
        out.append("")
        out.append(f"launch {n}: {sum(int(r[iS]) for r in body)} stall samples over {len(body)} SASS instructions")
        out.append("  " + ", ".join(f"{k[6:]}={sum(int(r[h.index(k)]) for r in body)}" for k in st if sum(int(r[h.index(k)]) for r in body)))
        cnt = collections.Counter()
        for r in body:
            s = r[iSrc].strip()
            if s.startswith('@'):
                s = s.split(None, 1)[1]
            cnt[s.split()[0].split('.')[0]] += int(r[iE])
        t = sum(cnt.values())
        out.append("  opcode mix: " + ", ".join(f"{k} {100 * v / t:.1f}%" for k, v in cnt.most_common(10)))
        for idx, r in sorted(enumerate(body), key=lambda x: -int(x[1][iS]))[:6]:
            out.append(f"  hot #{idx}: {r[iS]} samples  {r[iSrc].strip()[:72]}")
    with open(os.path.join(HERE, f"{tag}_stage_kernel.txt"), "w") as f:
        f.write("\n".join(out) + "\n")
    with open(os.path.join(HERE, "stage_kernel_traffic.json"), "w") as f:
        json.dump({"source": f"profiles/{tag}_stage_kernel.txt", "blocks_in_capture": nblocks, "dram_bytes_per_launch_in_capture": per_launch,
                   "dram_bytes_per_block_per_launch": per_launch / nblocks}, f, indent=1)
    print("\n".join(out[:30]))


if __name__ == "__main__":
    main()
