"""Decode the `stage clocks` lines printed by `STAGE_CLOCKS=1 python profiles/prof_infer.py <workload>` for the two-tiles-in-flight
kernel (csrc/infer_twin.cuh): per stage and tile, epilogue cycles (accumulator ready -> operands published) and the gap before it
(publish of the previous epilogue -> this tile's accumulator ready).    python profiles/decode_twin_clocks.py LOGFILE"""
import sys

NAMES = ["inproj"] + sum([[f"L{l}.att0", f"L{l}.att1", f"L{l}.ln1", f"L{l}.relu", f"L{l}.ln2"] for l in range(3)], []) + \
        ["dyn0", "dyn1", "nexth", "rL1", "rhead", "vL1", "vhead", "polh", "polout"]

for line in open(sys.argv[1]):
    if not line.startswith("stage clocks"):
        if line.strip():
            print(line.rstrip())
        continue
    t = eval(line.split(":", 1)[1])
    print(f"  gather {t[1] - t[0]} / {t[2] - t[1]} cycles; CTA total {t[-1]} cycles for two tiles")
    i, out, epi, gap = 3, [], 0, 0
    for n in NAMES:
        wa, pa, wb, pb = t[i:i + 4]
        out.append(f"{n} {pa - wa}/{pb - wb} (gap {wa - t[i - 1]}/{wb - pa})")
        epi += (pa - wa) + (pb - wb)
        gap += (wa - t[i - 1]) + (wb - pa)
        i += 4
    print("  " + "; ".join(out))
    print(f"  sum of epilogues {epi}, sum of gaps {gap}")
