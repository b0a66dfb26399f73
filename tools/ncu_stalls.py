"""Warp-stall summary of one kernel from the source page of an `ncu --set full --import-source on` capture.

    ncu -i <rep> --page source --csv --kernel-name regex:<name> --launch-count 1 > src.csv
    python tools/ncu_stalls.py src.csv <out.md> [window]

Writes the kernel-wide stall mix, the same per window of `window` consecutive SASS instructions (default 100) with
the opcodes that dominate the window (enough to tell the warp roles of a warp-specialised kernel apart), and the
instructions that collected the most samples."""
import collections
import csv
import sys


def num(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


def main():
    path, out = sys.argv[1], sys.argv[2]
    win = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    rows = list(csv.reader(open(path)))
    name = rows[0][1] if len(rows[0]) > 1 else "?"
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    seen, data = set(), []
    for r in rows[2:]:
        if len(r) < len(hdr) - 2 or r[ix["Address"]] in seen:
            continue
        seen.add(r[ix["Address"]])
        data.append(r)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(num(r[ix["# Samples"]]) for r in data)

    def opcode(r):
        t = r[ix["Source"]].split()
        return (t[1] if t and t[0].startswith("@") and len(t) > 1 else t[0] if t else "?").split(".")[0]

    with open(out, "w") as f:
        f.write(f"# Warp-stall samples: `{name}`\n\n{len(data)} SASS instructions, {int(tot)} samples.\n\n")
        f.write("| stall reason | samples | share |\n|---|---|---|\n")
        agg = {s: sum(num(r[ix[s]]) for r in data) for s in stalls}
        for s, v in sorted(agg.items(), key=lambda kv: -kv[1]):
            if v:
                f.write(f"| {s[6:]} | {int(v)} | {100 * v / tot:.1f} % |\n")
        f.write(f"\n## Windows of {win} instructions (>= 1 % of the samples)\n\n")
        f.write("| first address | samples | share | executed (warp-inst) | dominant opcodes | top stall reasons |\n|---|---|---|---|---|---|\n")
        for i in range(0, len(data), win):
            blk = data[i:i + win]
            s = sum(num(r[ix["# Samples"]]) for r in blk)
            if s < 0.01 * tot:
                continue
            ops = collections.Counter(opcode(r) for r in blk).most_common(4)
            st = sorted(((k[6:], sum(num(r[ix[k]]) for r in blk)) for k in stalls), key=lambda kv: -kv[1])[:4]
            ex = sum(num(r[ix["Instructions Executed"]]) for r in blk)
            f.write(f"| ...{blk[0][ix['Address']][-5:]} | {int(s)} | {100 * s / tot:.1f} % | {int(ex)} | "
                    f"{', '.join(f'{o} x{n}' for o, n in ops)} | {', '.join(f'{k} {int(v)}' for k, v in st)} |\n")
        f.write("\n## Instructions with the most samples\n\n| address | samples | dominant stall | instruction |\n|---|---|---|---|\n")
        for r in sorted(data, key=lambda r: -num(r[ix["# Samples"]]))[:25]:
            dom = max(stalls, key=lambda k: num(r[ix[k]]))
            f.write(f"| ...{r[ix['Address']][-5:]} | {int(num(r[ix['# Samples']]))} | {dom[6:]} | `{r[ix['Source']][:70]}` |\n")
    print(open(out).read()[:1500])


if __name__ == "__main__":
    main()
