"""Per-role instruction and stall budget of the fused phase-1 kernel from an `ncu --set full --import-source on`
report (source page, SASS view):

    ncu -i gpurun_out/prof_fwd_r1e.ncu-rep --page source --csv --print-source sass --kernel-name regex:k_qmlp > k1.csv
    python tools/sass_budget.py k1.csv <tiles in the launch>

Roles are recognised from the instruction stream itself (no hand-kept address ranges): a region is a maximal run
of SASS instructions with the same executed count (within 2 %), labelled by the opcodes it contains.
"""
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hdr = next(r for r in rows if "Instructions Executed" in r)
    data = [r for r in rows[rows.index(hdr) + 1:] if len(r) == len(hdr) and r[hdr.index("Instructions Executed")].isdigit()]
    # ncu prints the kernel once per source view; keep the first copy
    src = hdr.index("Source")
    first = data[0][src]
    for i in range(1, len(data)):
        if data[i][src] == first and i * 2 == len(data):
            data = data[:i]
            break
    return hdr, data


def label(ops):
    has = lambda *names: any(n in ops for n in names)
    if ops.get("SYNCS", 0) * 10 >= sum(ops.values()) and not has("UTCHMMA", "UTCQMMA", "UBLKCP"):
        return "mbarrier poll loop (try_wait, nanosleep, branch)"
    if has("UTCHMMA", "UTCQMMA"):
        return "mma issuer"
    if has("UBLKCP"):
        return "weight producer (bulk copy)"
    if has("MUFU") and has("STG"):
        return "epilogue: tanh + Q store"
    if has("FMNMX") and has("F2FP"):
        return "epilogue: bias + relu + split -> TMEM"
    if has("F2FP") and has("STS"):
        return "converter: split + STS (+ score FMAs)"
    if has("SHFL"):
        return "score row sums / arg-max keys"
    if has("LDG"):
        return "loads / addressing"
    return "other"


def main():
    path, tiles = sys.argv[1], int(sys.argv[2])
    hdr, data = load(path)
    iex, ismp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    stalls = [i for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
    ins = []
    for r in data:
        s = r[isrc].strip()
        toks = s.split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        ins.append((op, int(r[iex]), int(r[ismp]), [int(r[i] or 0) for i in stalls]))
    regions, start = [], 0
    for i in range(1, len(ins) + 1):
        if i == len(ins) or abs(ins[i][1] - ins[start][1]) > 0.02 * max(ins[start][1], 1) + 8:
            regions.append((start, i))
            start = i
    agg = {}
    for a, b in regions:
        ops = {}
        for op, ex, _, _ in ins[a:b]:
            ops[op] = ops.get(op, 0) + ex
        lab = label(ops) if sum(ops.values()) else "not executed"
        g = agg.setdefault(lab, {"ex": 0, "smp": 0, "st": [0] * len(stalls), "n": 0})
        for op, ex, smp, st in ins[a:b]:
            g["ex"] += ex
            g["smp"] += smp
            g["n"] += 1
            for j, v in enumerate(st):
                g["st"][j] += v
    tot_ex = sum(g["ex"] for g in agg.values())
    tot_smp = sum(g["smp"] for g in agg.values())
    print(f"SASS instructions: {len(ins)}; warp-instructions executed: {tot_ex} = {tot_ex / tiles:.0f} per 128-row tile "
          f"({tiles} tiles); stall samples: {tot_smp}")
    print(f"{'role':44s} {'static':>6s} {'instr/tile':>10s} {'share':>6s} {'samples':>8s}  top stall reasons")
    for lab, g in sorted(agg.items(), key=lambda kv: -kv[1]["ex"]):
        top = sorted(zip((hdr[i][6:] for i in stalls), g["st"]), key=lambda kv: -kv[1])[:4]
        print(f"{lab:44s} {g['n']:6d} {g['ex'] / tiles:10.0f} {100 * g['ex'] / tot_ex:5.1f}% {g['smp']:8d}  "
              + ", ".join(f"{k} {v}" for k, v in top if v))


if __name__ == "__main__":
    main()
