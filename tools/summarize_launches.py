"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares of ONE steady-state step.

usage: python tools/summarize_launches.py gpurun_out/launches.csv [out.json]
A step is delimited by the loss kernel (bq::ce_mean_kernel, or torch's nll_loss_forward) that ends every forward; the last complete step is used
(step 1 holds the one-off PTQ weight quantisation and is never chosen when a later one exists).
ncu's per-launch times are cold-cache and serialised: compare SHARES with bench.py's event-timed shares, not absolutes."""
import csv
import json
import re
import sys


def short(name: str) -> str:
    name = re.sub(r"^void ", "", name)
    m = re.match(r"(?:bq::)?(\w+)(<[^(]*>)?\(", name)
    if m and ("bq::" in name or m.group(1) in ("quant_rows_kernel", "gemm_bf16_tn_kernel", "attention_causal_kernel")):
        return "bq::" + m.group(1) + (m.group(2) or "")
    m = re.search(r"at::native::(?:\(anonymous namespace\)::|<unnamed>::)?(\w+)", name)
    if m:
        extra = re.search(r"(launch_clamp_scalar|CUDAFunctor_add|direct_copy|MulFunctor|FillFunctor|BinaryFunctor)", name)
        return "at::" + m.group(1) + ("[" + extra.group(1) + "]" if extra else "")
    return name[:80]


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        rows.append((int(r["ID"]), r["Kernel Name"], us, r["Grid Size"], r["Block Size"]))
    ends = [i for i, r in enumerate(rows) if "nll_loss_forward_reduce" in r[1] or "ce_mean_kernel" in r[1]]
    if len(ends) >= 2:
        lo, hi = ends[-2] + 1, ends[-1] + 1
        which = f"launches {rows[lo][0]}..{rows[hi - 1][0]} (step {len(ends)} of {len(ends)} seen)"
    else:
        lo, hi = 0, len(rows)
        which = "all captured launches (no step boundary found)"
    step = rows[lo:hi]
    agg = {}
    for _, name, us, grid, block in step:
        k = short(name)
        a = agg.setdefault(k, {"launches": 0, "total_us": 0.0, "max_us": 0.0})
        a["launches"] += 1
        a["total_us"] += us
        a["max_us"] = max(a["max_us"], us)
    tot = sum(a["total_us"] for a in agg.values())
    out = {"source": path, "window": which, "launches_in_step": len(step), "sum_of_kernel_time_ms": tot / 1e3,
           "note": "ncu serialises launches and times each one cold; shares are comparable with bench.py, absolutes are not",
           "kernels": []}
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["total_us"]):
        out["kernels"].append({"kernel": k, "launches": a["launches"], "total_ms": round(a["total_us"] / 1e3, 3),
                               "avg_us": round(a["total_us"] / a["launches"], 2), "max_us": round(a["max_us"], 2),
                               "share": round(a["total_us"] / tot, 4)})
    txt = json.dumps(out, indent=1)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(txt + "\n")
    print(txt)


if __name__ == "__main__":
    main()
