"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one document's launches grouped by kernel."""
import collections, csv, re, sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    seq = []
    for x in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", x["Kernel Name"]).replace("void ", "").replace("dvd::", "")
        try:
            t = float(x["Metric Value"])
        except ValueError:
            continue
        u = x["Metric Unit"]
        t = t / 1000 if u == "ns" else (t * 1000 if u == "ms" else t)
        seq.append((name, x.get("Grid Size", ""), t))
    idx = [i for i, (n, g, t) in enumerate(seq) if n.startswith("k_unwarp")]
    run = seq[idx[-2] + 1: idx[-1] + 1] if len(idx) >= 2 else seq
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, g, t in run:
        agg[n][0] += 1
        agg[n][1] += t
    tot = sum(v[1] for v in agg.values())
    print(f"# one document (static pass + 3 DDIM steps x 2 hypotheses + unwarp): {len(run)} launches, {tot:.1f} us summed device time")
    print("# (ncu serialises launches and runs them cold: compare SHARES, not absolutes)")
    print(f"{'kernel':46s} {'n':>4s} {'total_us':>10s} {'avg_us':>8s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:46]:46s} {v[0]:4d} {v[1]:10.1f} {v[1] / v[0]:8.1f} {100 * v[1] / tot:6.1f}%")
    agg2 = collections.defaultdict(lambda: [0, 0.0])
    for n, g, t in run:
        if "gemm" in n or "attn" in n:
            agg2[(n, g)][0] += 1
            agg2[(n, g)][1] += t
    print("\n# tensor-core kernels by grid")
    for k, v in sorted(agg2.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[0][:30]:30s} {k[1]:>16s} n={v[0]:3d} total={v[1]:8.1f} us avg={v[1] / v[0]:7.1f} us")


if __name__ == "__main__":
    main(sys.argv[1])
