"""Turn ncu exports into the committed summaries under profiles/.

    python tools/ncu_summary.py launches <launches.csv> <out.md>      # --metrics gpu__time_duration.sum pass
    python tools/ncu_summary.py full <raw.csv> <out.md> <out.json>    # `ncu -i rep --page raw --csv` of a --set full capture
"""
import collections
import csv
import json
import re
import sys


def short(name: str) -> str:
    name = re.sub(r"^void\s+", "", name)
    name = name.replace("nrx::", "").replace("(int)", "").replace("(bool)", "")
    return re.sub(r"\(.*\)$", "", name)


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    i_name, i_val = hdr.index("Kernel Name"), hdr.index("Metric Value")
    i_unit = hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows:
        if r is hdr or r[0] == "ID" or not r[0].strip('"').isdigit():
            continue
        v = float(r[i_val].replace(",", ""))
        v = v / 1e3 if r[i_unit] in ("nsecond", "ns") else v * 1e3 if r[i_unit] in ("msecond", "ms") else v
        a = agg.setdefault(short(r[i_name]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write("| kernel | launches | avg us | total us | share |\n|---|---|---|---|---|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {t / n:.1f} | {t:.0f} | {100 * t / tot:.1f}% |\n")
        f.write(f"\nTotal {tot:.0f} us over {sum(v[0] for v in agg.values())} launches.\n")
    print(open(out).read())


METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma_type_fp16.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__grid_size"]


def full(path, out_md, out_json):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rows[2:]:
        d = {"kernel": short(r[col["Kernel Name"]])}
        for m in METRICS:
            if m in col:
                v = r[col[m]].replace(",", "")
                try:
                    d[m] = float(v)
                except ValueError:
                    d[m] = v
                d[m + "|unit"] = units[col[m]]
        res.append(d)
    def to_bytes(d, m):
        u = d.get(m + "|unit", "byte")
        return d.get(m, 0.0) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    js = {}
    with open(out_md, "w") as f:
        f.write("| kernel | us | DRAM read MB | DRAM write MB | tensor pipe % | FMA pipe % | fp16 FMA inst % | issue % | smem wavefront % | regs |\n")
        f.write("|---|---|---|---|---|---|---|---|---|---|\n")
        for d in res:
            rd, wr = to_bytes(d, "dram__bytes_read.sum"), to_bytes(d, "dram__bytes_write.sum")
            f.write("| `{}` | {:.1f} | {:.1f} | {:.1f} | {:.1f} | {:.1f} | {:.1f} | {:.1f} | {:.1f} | {:.0f} |\n".format(
                d["kernel"], d.get("gpu__time_duration.sum", 0), rd / 1e6, wr / 1e6,
                d.get(METRICS[3], 0), d.get(METRICS[4], 0), d.get(METRICS[5], 0), d.get(METRICS[6], 0),
                d.get(METRICS[7], 0), d.get("launch__registers_per_thread", 0)))
            js[d["kernel"]] = {"dram_bytes_per_launch": rd + wr, "us": d.get("gpu__time_duration.sum", 0)}
    # keyed to the kernel sources the capture was taken on: bench.py only reports these traffic figures while the
    # sources are unchanged (same hash as bench.source_hash())
    import hashlib, os
    h = hashlib.sha256()
    csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "neural_rx_b200", "csrc")
    for name in sorted(os.listdir(csrc)):
        with open(os.path.join(csrc, name), "rb") as f:
            h.update(name.encode() + b"\0" + f.read())
    merged = {"source_hash": h.hexdigest()[:16], "kernels": js}
    if os.path.exists(out_json):
        try:
            old = json.load(open(out_json))
            if old.get("source_hash") == merged["source_hash"]:
                old["kernels"].update(js)
                merged = old
        except Exception:
            pass
    json.dump(merged, open(out_json, "w"), indent=1)
    print(open(out_md).read())


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4])
