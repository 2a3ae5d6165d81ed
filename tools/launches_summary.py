"""Summarise the ncu launch list of one training step (gpurun_out/launches.csv from tools/gpu_round.sh):
per-kernel totals -> profiles/<tag>_ncu_launches_step_summary.txt, compact list -> profiles/<tag>_ncu_launches_step.csv and,
when the dram metrics were collected, the average DRAM traffic per conv_igemm launch -> profiles/ncu_conv_traffic.json
(bench.py reports it as roofline.traffic).
    python tools/launches_summary.py gpurun_out/launches.csv r01"""
import collections
import csv
import json
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def short(n: str) -> str:
    return re.sub(r"\(.*", "", n).replace("void ", "")[:70]


def main(path: str, tag: str) -> None:
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    gi, bi = hdr.index("Grid Size"), hdr.index("Block Size")
    launches = collections.OrderedDict()
    for r in data:
        d = launches.setdefault(r[0], {"kernel": r[ki], "grid": r[gi], "block": r[bi]})
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6,
                 "Gbyte": 1e9, "nsecond": 1e-3}.get(u, 1.0)
        d[r[mi]] = v * scale
    ids = list(launches)
    # first kernel of a training step: the batched weight re-pack (older captures: the layout conversion)
    starts = [i for i, k in enumerate(ids) if "pack_weights_batched" in launches[k]["kernel"]]
    if not starts:
        starts = [i for i, k in enumerate(ids) if "nchw_to_nhwc" in launches[k]["kernel"]]
    step = ids[starts[-1]:] if starts else ids
    # native optimizer (round 2): a step ENDS with the fused SGD + re-pack launch followed by vtb_sgd_step
    ends = [i for i, k in enumerate(ids) if "sgd_step_kernel" in launches[k]["kernel"]]
    if len(ends) >= 2:
        step = ids[ends[-2] + 1: ends[-1] + 1]
    agg, tot = collections.OrderedDict(), 0.0
    conv_bytes, conv_n = 0.0, 0
    have_dram = False
    for k in step:
        d = launches[k]
        t = d.get("gpu__time_duration.sum", 0.0)
        a = agg.setdefault(short(d["kernel"]), [0, 0.0, 0.0])
        by = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        have_dram |= "dram__bytes_read.sum" in d
        a[0] += 1; a[1] += t; a[2] += by; tot += t
        if "conv_igemm_kernel" in d["kernel"]:
            conv_bytes += by; conv_n += 1
    out = ROOT / "profiles" / f"{tag}_ncu_launches_step_summary.txt"
    with open(out, "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum[,dram__bytes_*] --clock-control none : python tools/one_step.py "
                "cspdarknet53 256 176 2  (second step only; serialised, cold-cache: compare SHARES)\n")
        f.write(f"total {tot / 1e3:.3f} ms over {len(step)} launches\n")
        for k, (n, t, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            extra = f"  dram {by / 1e6:9.1f} MB" if have_dram else ""
            f.write(f"{t / 1e3:9.3f} ms {100 * t / tot:5.1f}%  n={n:4d}  {k}{extra}\n")
    with open(ROOT / "profiles" / f"{tag}_ncu_launches_step.csv", "w") as f:
        f.write("id,kernel,grid,block,gpu__time_duration.sum_us,dram_bytes\n")
        for k in step:
            d = launches[k]
            by = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
            f.write(f'{k},"{short(d["kernel"])[:60]}","{d["grid"]}","{d["block"]}",{d.get("gpu__time_duration.sum", 0.0):.2f},{by:.0f}\n')
    if have_dram and conv_n:
        (ROOT / "profiles" / "ncu_conv_traffic.json").write_text(json.dumps(
            {"kernel": "conv_igemm_kernel (all launches of one cspdarknet53 bs256 176px step)", "launches": conv_n,
             "avg_dram_bytes_per_launch": conv_bytes / conv_n, "source": f"profiles/{tag}_ncu_launches_step.csv"}))
    print(open(out).read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "r01")
