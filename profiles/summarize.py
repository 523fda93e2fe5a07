#!/usr/bin/env python
"""Turn an `ncu --set full --import-source on` report of the step kernels into the tracked
summaries under profiles/ (the .ncu-rep itself stays in gpurun_out/, which is scratch).

    python profiles/summarize.py gpurun_out/r02_full.ncu-rep r02 --worlds 8192

The report may hold one launch of the fused kernel (pxStepKernel<NT, 0>) or one launch each of the
split pipeline's sweep kernel (pxStepKernel<NT, 1>) and solver kernel (pxSolveKernel); every launch
covers ONE substep of every world in the split pipeline, `--fused` substeps in the fused one.
Writes
    profiles/<tag>_ncu_raw_<kernel>.csv     every raw metric of the launch (metric,unit,value)
    profiles/<tag>_segments_<kernel>.txt    SASS segments between barriers: share of samples and of
                                            instructions, lanes per instruction, shared-memory wavefronts
                                            and global requests per world, stall mix
    profiles/<tag>_hot_instructions_<kernel>.txt  the 40 most-sampled SASS instructions with their stalls
    profiles/<tag>_kernel_facts.json        the numbers bench.py folds into its JSON line
Runs here (no GPU): needs only the `ncu` CLI to read the report.
"""
from __future__ import annotations

import argparse
import collections
import csv
import io
import json
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def ncu_csv(rep: str, *args: str):
    out = subprocess.run(["ncu", "-i", rep, "--csv", *args], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def to_bytes(value: str, unit: str) -> float:
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return float(value) * scale


def kernel_key(name: str) -> str:
    if "pxSolveKernel" in name or "pxReplayKernel" in name:
        return "solve"
    m = re.search(r"pxStepKernel<[^>]*?,\s*(?:\(int\))?(\d)\s*>", name)
    if m:
        return "sweep" if m.group(1) == "1" else "fused"
    return name.split("(")[0].strip()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("tag")
    ap.add_argument("--worlds", type=int, required=True)
    ap.add_argument("--fused", type=int, default=5, help="substeps per launch of the FUSED kernel (split launches cover one)")
    a = ap.parse_args()

    raw = ncu_csv(a.report, "--page", "raw")
    hdr, units = raw[0], raw[1]
    facts = {"worlds": a.worlds, "source": os.path.basename(a.report), "kernels": {}}
    seen = set()
    for vals in raw[2:]:
        if len(vals) != len(hdr):
            continue
        metrics = {h: (units[i], vals[i]) for i, h in enumerate(hdr)}
        key = kernel_key(metrics["Kernel Name"][1])
        if key in seen:
            continue
        seen.add(key)
        substeps = a.fused if key == "fused" else 1
        world_substeps = a.worlds * substeps
        with open(os.path.join(HERE, f"{a.tag}_ncu_raw_{key}.csv"), "w") as f:
            f.write("metric,unit,value\n")
            for h in sorted(metrics):
                if "__" in h or h.startswith("launch"):
                    f.write(f"{h},{metrics[h][0]},{metrics[h][1]}\n")

        def m(name):
            return float(metrics[name][1])

        rd = to_bytes(metrics["dram__bytes_read.sum"][1], metrics["dram__bytes_read.sum"][0])
        wr = to_bytes(metrics["dram__bytes_write.sum"][1], metrics["dram__bytes_write.sum"][0])
        dur_unit = metrics["gpu__time_duration.sum"][0]
        dur_ms = m("gpu__time_duration.sum") * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(dur_unit, 1.0)
        facts["kernels"][key] = {
            "kernel": metrics["Kernel Name"][1].strip(),
            "substeps_per_launch": substeps,
            "gpu_time_ms_under_ncu": dur_ms,
            "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
            "warp_instructions_per_launch": m("smsp__inst_executed.sum"),
            "warp_instructions_per_world_substep": m("smsp__inst_executed.sum") / world_substeps,
            "issue_active_pct": m("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "lanes_per_instruction": m("smsp__thread_inst_executed_per_inst_executed.ratio"),
            "registers_per_thread": m("launch__registers_per_thread"),
            "shared_mem_per_block_bytes": m("launch__shared_mem_per_block_dynamic") * 1e3,
            "l1_hit_rate_pct": m("l1tex__t_sector_hit_rate.pct"),
            "l2_hit_rate_pct": m("lts__t_sector_hit_rate.pct"),
            # pipe utilisation, % of peak: ALU, the FMA pipe (the non-fused FMUL / FADD issue there), LSU
            # instruction issue, and the L1 / shared-memory DATA pipe in wavefronts -- the busiest unit
            "pipe_alu_pct": m("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
            "pipe_fma_pct": m("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
            "pipe_lsu_pct": m("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
            "l1_data_pipe_wavefronts_pct": m("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
            "shared_wavefronts_pct": m("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
            "shared_bank_conflict_wavefronts": m("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
            "shared_wavefronts": m("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
            "global_load_wavefronts": m("l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum"),
            "achieved_warps_pct_of_max": m("sm__warps_active.avg.pct_of_peak_sustained_active"),
            "eligible_warps_per_scheduler_cycle": m("smsp__warps_eligible.avg.per_cycle_active"),
        }

    # ---- per-instruction view, one section per kernel
    src = ncu_csv(a.report, "--page", "source")
    sections, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            sections.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    done = set()
    for sec in sections:
        key = kernel_key(sec["name"])
        if key in done or not sec["rows"]:
            continue
        done.add(key)
        hdr = sec["rows"][0]
        data = [r for r in sec["rows"][1:] if len(r) == len(hdr)]
        ix = {h: i for i, h in enumerate(hdr)}
        i_src, i_smp, i_ie, i_te = (ix[x] for x in ("Source", "# Samples", "Instructions Executed", "Thread Instructions Executed"))
        i_ws, i_wi, i_g = ix["L1 Wavefronts Shared"], ix["L1 Wavefronts Shared Ideal"], ix["L1 Tag Requests Global"]
        stalls = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        total_smp = sum(int(r[i_smp]) for r in data) or 1
        total_ie = sum(int(r[i_ie]) for r in data) or 1
        worlds = float(a.worlds)
        bars = [k for k, r in enumerate(data) if "BAR.SYNC" in r[i_src] or "BAR.RED" in r[i_src]]
        with open(os.path.join(HERE, f"{a.tag}_segments_{key}.txt"), "w") as f:
            f.write(f"# SASS segments between CTA barriers of {sec['name'].strip()} ({len(data)} instructions); samples attributed\n"
                    f"# to a barrier wait show up on the first instruction AFTER the barrier.  {os.path.basename(a.report)}\n"
                    f"# columns: first-last SASS index, share of warp samples, share of executed warp instructions, warp\n"
                    f"#          instructions per world (one launch), lanes per instruction, shared-memory wavefronts per world\n"
                    f"#          (and the conflict-free ideal), global-memory requests per world, top stall reasons (% of segment)\n")
            prev = 0
            for b in bars + [len(data) - 1]:
                seg = data[prev:b + 1]
                smp = sum(int(r[i_smp]) for r in seg)
                ie = sum(int(r[i_ie]) for r in seg)
                te = sum(int(r[i_te]) for r in seg)
                ws = sum(int(r[i_ws]) for r in seg)
                wi = sum(int(r[i_wi]) for r in seg)
                g = sum(int(r[i_g]) for r in seg)
                if smp > 0.004 * total_smp or ie > 0.004 * total_ie:
                    c = collections.Counter()
                    for r in seg:
                        for i, name in stalls:
                            c[name] += int(r[i])
                    tot = sum(c.values()) or 1
                    mix = " ".join(f"{k}:{100 * v / tot:.0f}" for k, v in c.most_common(5))
                    f.write(f"{prev:6d}-{b:6d}  samples {100 * smp / total_smp:5.1f}%  instr {100 * ie / total_ie:5.1f}%  "
                            f"{ie / worlds:8.0f}/w  lanes {te / max(ie, 1):4.1f}  smem-wf {ws / worlds:7.0f} (ideal {wi / worlds:6.0f})  "
                            f"global-req {g / worlds:6.0f}  {mix}\n")
                prev = b + 1
        with open(os.path.join(HERE, f"{a.tag}_hot_instructions_{key}.txt"), "w") as f:
            f.write("# most-sampled SASS instructions: index, executions, share of samples, top stalls, instruction\n")
            for k in sorted(range(len(data)), key=lambda k: -int(data[k][i_smp]))[:40]:
                r = data[k]
                st = sorted(((int(r[i]), name) for i, name in stalls), reverse=True)[:3]
                f.write(f"{k:6d} {int(r[i_ie]):11d} {100 * int(r[i_smp]) / total_smp:5.2f}%  "
                        + " ".join(f"{name}:{v}" for v, name in st if v) + "  | " + r[i_src].strip()[:90] + "\n")

    with open(os.path.join(HERE, f"{a.tag}_kernel_facts.json"), "w") as f:
        json.dump(facts, f, indent=1)
        f.write("\n")
    print(json.dumps({k: {x: v[x] for x in ("gpu_time_ms_under_ncu", "warp_instructions_per_world_substep", "issue_active_pct",
                                            "l1_data_pipe_wavefronts_pct", "dram_bytes_per_launch")} for k, v in facts["kernels"].items()}))


if __name__ == "__main__":
    main()
