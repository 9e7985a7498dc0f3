#!/usr/bin/env python
"""Turn an .ncu-rep (brought back from the GPU box in gpurun_out/) into the small, committed summary
the bench and the reviewers read:  python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/NAME [POINTS]
(POINTS = query points of the profiled launch: adds the per-point counts bench.py's roofline line is built from)

Writes NAME_summary.json (key counters of the first profiled launch of every kernel) and
NAME_hot_sass.txt (the 40 SASS lines with the most stall samples, needs -lineinfo + --import-source).
Runs `ncu -i` locally; no GPU needed."""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration_ms",
    "launch__grid_size": "grid", "launch__block_size": "block", "launch__registers_per_thread": "registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slots_busy_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "pipe_fp64_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "pipe_xu_pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1_lsu_wavefronts_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_wavefronts_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct", "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    # raw counts behind the roofline line of bench.py (round-1 verdict item 2): everything below is a COUNT per launch
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum": "smem_wavefronts_ld",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum": "smem_wavefronts_st",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "smsp__sass_inst_executed_op_shared_ld.sum": "smem_ld_warp_instructions",
    "smsp__sass_inst_executed_op_shared_st.sum": "smem_st_warp_instructions",
    "sm__cycles_elapsed.avg": "sm_cycles_elapsed",
    "sm__cycles_elapsed.avg.per_second": "sm_clock_ghz",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed": "dfma_thread_inst_per_cycle",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed": "dadd_thread_inst_per_cycle",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed": "dmul_thread_inst_per_cycle",
    "sm__inst_executed_pipe_fp64.sum": "pipe_fp64_warp_instructions",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum": "dfma_thread_inst",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum": "dadd_thread_inst",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum": "dmul_thread_inst",
    "launch__sm_count": "sm_count",
}
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, name = sys.argv[1], sys.argv[2]
    points = int(float(sys.argv[3])) if len(sys.argv) > 3 else None   # points per profiled launch (for per-point figures)
    rows = ncu_csv(rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    kcol = hdr.index("Kernel Name")
    out = {"report": rep, "kernels": []}
    seen = set()
    for r in data:
        if r[kcol] in seen:
            continue
        seen.add(r[kcol])
        k = {"kernel": r[kcol]}
        for i, h in enumerate(hdr):
            if h in KEYS and r[i] not in ("", "nan", "-nan"):
                v = float(r[i].replace(",", ""))
                if KEYS[h] in ("dram_read", "dram_write", "duration_ms"):
                    v *= UNIT_SCALE.get(units[i], 1.0)
                k[KEYS[h]] = v
        if "dram_read" in k and "dram_write" in k:
            k["dram_bytes_per_launch"] = k["dram_read"] + k["dram_write"]
        cyc = k.get("sm_cycles_elapsed")
        if cyc:  # thread-instruction counts from the per-cycle rates (what --set full records) when no .sum was captured
            for op in ("dfma", "dadd", "dmul"):
                if op + "_thread_inst" not in k and op + "_thread_inst_per_cycle" in k:
                    k[op + "_thread_inst"] = k[op + "_thread_inst_per_cycle"] * cyc
            if all(op + "_thread_inst" in k for op in ("dfma", "dadd", "dmul")):
                k["fp64_thread_inst"] = k["dfma_thread_inst"] + k["dadd_thread_inst"] + k["dmul_thread_inst"]
            sms = k.get("sm_count") or k.get("grid") or 148
            if "smem_wavefronts" in k:  # 1 wavefront per clock per SM is the pipe's peak
                k["smem_wavefront_peak_per_launch"] = cyc * sms
                k["smem_wavefront_frac"] = k["smem_wavefronts"] / (cyc * sms)
            if "fp64_thread_inst" in k:  # 64 FP64 lanes per clock per SM (measured: asg_measure_fp64_peak)
                k["fp64_lane_frac"] = k["fp64_thread_inst"] / (cyc * sms * 64.0)
        out["kernels"].append(k)
    if out["kernels"]:
        out["dram_bytes_per_launch"] = out["kernels"][0].get("dram_bytes_per_launch")
    if points:
        out["points_per_launch"] = points
        for k in out["kernels"]:
            for key in ("smem_wavefronts", "smem_ld_warp_instructions", "fp64_thread_inst", "warp_instructions"):
                if key in k:
                    k[key + "_per_point"] = k[key] / points
    json.dump(out, open(name + "_summary.json", "w"), indent=1)
    try:
        src = ncu_csv(rep, "source")
        h = src[1]
        ia, isrc, ismp = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
        first = src[2:]
        for n, r in enumerate(first):  # the page lists every profiled launch in turn: keep the first kernel
            if r and r[0] == "Kernel Name":
                first = first[:n]
                break
        body = [(int(r[ismp] or 0), int(r[ia] or 0), r[isrc].strip()) for r in first if len(r) > ismp and r[0] != "Address"]
        tot_s, tot_i = sum(b[0] for b in body) or 1, sum(b[1] for b in body) or 1
        with open(name + "_hot_sass.txt", "w") as f:
            f.write(f"# {src[0][1] if len(src[0]) > 1 else ''}\n# total warp instructions {tot_i}, stall samples {tot_s}\n")
            f.write("# samples%  instr%  SASS\n")
            for s, i, t in sorted(body, reverse=True)[:40]:
                f.write(f"{100 * s / tot_s:7.2f} {100 * i / tot_i:7.2f}  {t}\n")
    except Exception as e:  # source page is optional
        print("no source page:", e)
    print(json.dumps(out["kernels"], indent=1))


if __name__ == "__main__":
    main()
