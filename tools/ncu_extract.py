"""Extract the roofline-relevant metrics of every captured launch from an .ncu-rep (`ncu --set full`) into JSON.

usage: python tools/ncu_extract.py gpurun_out/prof.ncu-rep profiles/rNN_ncu_<what>.json
Runs `ncu -i <rep> --page raw --csv` (works without a GPU)."""
import csv
import io
import json
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_bytes.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_lsu.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_sleeping_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct", "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct",
        "smsp__warp_issue_stalled_membar_per_warp_active.pct", "smsp__warp_issue_stalled_tex_throttle_per_warp_active.pct"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    if rep.endswith(".csv"):                 # already exported on the GPU box (`ncu -i rep --page raw --csv`)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw[raw.index('"ID"'):])))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")][:120]}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                try:
                    d[w] = float(r[i].replace(",", ""))
                except ValueError:
                    d[w] = r[i]
                d[w + "__unit"] = units[i]
        res.append(d)
    # compact: drop the unit keys into one map
    unitmap = {}
    for d in res:
        for k in [k for k in d if k.endswith("__unit")]:
            unitmap[k[:-6]] = d.pop(k)
    json.dump({"source": rep, "units": unitmap, "launches": res}, open(out, "w"), indent=1)
    for d in res:
        t = d.get("gpu__time_duration.sum", 0)
        print(f'{d["kernel"][:70]:70s} {t:10.1f} {unitmap.get("gpu__time_duration.sum")}  dramR {d.get("dram__bytes_read.sum")} W {d.get("dram__bytes_write.sum")} '
              f'tensor% {d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")} issue% {d.get("smsp__issue_active.avg.pct_of_peak_sustained_active")}')


if __name__ == "__main__":
    main()
