#!/usr/bin/env python
"""Prints the few ncu metrics this project tracks from a .ncu-rep (one line block per kernel launch);
with a second argument also writes them as JSON (the file bench.py reads its `roofline.executed` block from)."""
import csv
import subprocess
import sys

WANT = ['Grid Size', 'Block Size', 'gpu__time_duration.sum', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'smsp__sass_inst_executed_op_shared_ld.sum', 'smsp__sass_inst_executed_op_shared_st.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']


def main(path, json_out=None):
    import json
    records = []
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("==", d.get("Kernel Name", "?")[:100])
        for k in WANT:
            if k in d:
                print(f"  {k} = {d[k]} {units[hdr.index(k)]}")
        st = {k.replace('smsp__pcsamp_warps_issue_stalled_', ''): int(d[k]) for k in hdr
              if 'pcsamp_warps_issue_stalled' in k and 'not_issued' not in k and d[k].isdigit()}
        tot = sum(st.values()) or 1
        print("  stalls:", ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
        rec = {"kernel": d.get("Kernel Name", "?"), "source": path}
        for k in WANT:
            if k in d:
                try:
                    rec[k] = float(d[k].replace(",", ""))
                except ValueError:
                    rec[k] = d[k]
                rec.setdefault("units", {})[k] = units[hdr.index(k)]
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[hdr.index(k)])
                if scale and isinstance(rec[k], float):
                    rec[k + "_bytes"] = rec[k] * scale
        rec["stalls_pct"] = {k: round(100 * v / tot, 2) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:10]}
        records.append(rec)
    if json_out:
        with open(json_out, "w") as f:
            json.dump(records, f, indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
