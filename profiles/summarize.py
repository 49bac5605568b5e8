#!/usr/bin/env python
"""Turn the raw ncu outputs under gpurun_out/ into the committed summaries under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r1.csv  > profiles/r1_launches.md
    python profiles/summarize.py kernel  gpurun_out/prof_r1_gemm.ncu-rep gemm_topk > profiles/r1_gemm_topk.md
    python profiles/summarize.py traffic gemm_topk_kernel:gpurun_out/prof_r2_gemm_long.ncu-rep:10000000:512:4096:100 \
                                         gemm_topk_kernel:gpurun_out/prof_r2_gemm_short.ncu-rep:1250000:512:4096:100 \
                                         scan_topk_kernel:gpurun_out/prof_r2_scan.ncu-rep:10000000:512:1:10 > profiles/traffic.json
"""
import collections
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'sm__cycles_elapsed.avg', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic']


def launches(path):
    lines = open(path).read().splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    agg, seq = collections.OrderedDict(), []
    for row in csv.DictReader(lines[start:]):
        name = row['Kernel Name'].split('(')[0][-70:]
        val = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        val = val / 1e3 if unit == 'ns' else (val * 1e3 if unit == 'ms' else val)
        seq.append((name, val))
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += val
    total = sum(v for _, v in seq)
    print(f"ncu launch list ({len(seq)} launches, {total / 1e3:.2f} ms of device time; cold-cache and serialised:"
          " compare SHARES, not absolutes)\n")
    print("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.1f} | {100 * v[1] / total:.1f}% |")
    idx = [i for i, (n, _) in enumerate(seq) if 'gemm_topk' in n]
    if idx:
        print("\nOne search+rerank step (launch order around the last GEMM):\n")
        print("| us | kernel |\n|---:|---|")
        step = seq[idx[-1] - 1: idx[-1] + 4]
        st = sum(v for _, v in step)
        for n, v in step:
            print(f"| {v:.1f} | `{n}` |")
        print(f"\nGEMM share of the step: {100 * seq[idx[-1]][1] / st:.1f}%")


def kernel(rep, pat, top=25):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        if pat not in r[idx['Kernel Name']]:
            continue
        print(f"## `{r[idx['Kernel Name']][:90]}`\n\n| metric | value | unit |\n|---|---:|---|")
        for w in KEYS:
            if w in idx:
                print(f"| {w} | {r[idx[w]]} | {units[idx[w]]} |")
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + pat],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {'hdr': None, 'rows': []}; blocks.append(cur); continue
        if cur is not None and cur['hdr'] is None:
            cur['hdr'] = r; continue
        if cur is not None:
            cur['rows'].append(r)
    b = blocks[0]
    h = {n: i for i, n in enumerate(b['hdr'])}
    stalls = [n for n in b['hdr'] if n.startswith('stall_') and 'Not Issued' not in n]
    tot, lines = 0, []
    for r in b['rows']:
        try:
            s = int(r[h['# Samples']])
        except Exception:
            continue
        tot += s; lines.append((s, r))
    agg = {n: 0 for n in stalls}
    for s, r in lines:
        for n in stalls:
            try:
                agg[n] += int(r[h[n]])
            except Exception:
                pass
    print(f"\nWarp-stall samples: {tot}\n\n| reason | samples | share |\n|---|---:|---:|")
    for n, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
        print(f"| {n} | {v} | {100 * v / max(tot, 1):.1f}% |")
    lines.sort(key=lambda x: -x[0])
    print(f"\nTop {top} SASS instructions by samples:\n\n| samples | share | executed | instruction | top stall |\n|---:|---:|---:|---|---|")
    for s, r in lines[:top]:
        t = sorted(((int(r[h[n]] or 0), n) for n in stalls), reverse=True)[0]
        print(f"| {s} | {100 * s / tot:.1f}% | {r[h['Instructions Executed']]} | `{r[h['Source']][:70]}` | {t[1]} |")


def traffic(specs):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch per kernel -> the table bench.py reads
    (roofline.traffic when the workload matches)."""
    import json
    out = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` captures "
                       "(profiles/summarize.py traffic); bench.py reports these as roofline.traffic when the workload matches"}
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for spec in specs:
        name, rep, rows, dim, batch, k = spec.split(":")
        raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rws = list(csv.reader(raw.splitlines()))
        hdr, units = rws[0], rws[1]
        idx = {h: i for i, h in enumerate(hdr)}
        r = [x for x in rws[2:] if name.split("_kernel")[0] in x[idx['Kernel Name']]][0]
        tot = sum(float(r[idx[m]].replace(',', '')) * mult[units[idx[m]]] for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
        out.setdefault(name, []).append({"rows": int(rows), "dim": int(dim), "batch": int(batch), "k": int(k),
                                         "bytes": int(tot), "kernel": r[idx['Kernel Name']][:80],
                                         "source": rep.split('/')[-1]})
    print(json.dumps(out, indent=2))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    elif sys.argv[1] == "traffic":
        traffic(sys.argv[2:])
    else:
        kernel(sys.argv[2], sys.argv[3])
