#!/usr/bin/env python3
"""Loop-carried dependence depth of a SASS loop (static estimate, no GPU needed).

usage: sass_depth.py <cubin-or-.o-or-.so> <kernel-substring> <loop-start-addr> <loop-end-addr> [skip-ranges...]
  addresses in hex as cuobjdump prints them (the loop's first instruction and its closing BRA); optional skip ranges
  `lo-hi` name rarely executed blocks (housekeeping, the once-per-lane finish) to leave out.

The loop body is unrolled twice; every instruction gets the earliest issue time allowed by its register / predicate /
uniform-register inputs (producer issue time + producer latency) and by in-order issue (one instruction per clock).  The
time between the two copies of the last instruction is the steady-state length of one trip for a single warp that owns
its scheduler.  Latencies are round numbers from the Blackwell notes (ALU/FMA 5 clocks to a dependent use, predicates 6,
shared-memory loads 30, POPC 8, global loads 300); the point is which chain is longest, not the third digit.
"""
import re
import subprocess
import sys

LAT = {"LDS": 30, "LDG": 300, "LDC": 20, "POPC": 8, "ISETP": 6, "PLOP3": 6, "LOP3": 5, "SEL": 5, "IMAD": 5, "IADD3": 5,
       "VIADD": 5, "SHF": 5, "PRMT": 5, "VIMNMX": 5, "STS": 1, "ATOMG": 300, "BRA": 1, "BSSY": 1, "BSYNC": 1, "S2UR": 20,
       "ULOP3": 5, "UIADD3": 5, "UMOV": 5, "ULEA": 5, "MOV": 5, "P2R": 6, "R2P": 6, "VOTE": 6, "LEA": 5, "I2FP": 6, "F2I": 8,
       "FMUL": 5, "FADD": 5, "FFMA": 5}


def parse(line):
    m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)\s*(.*?);", line)
    if not m:
        return None
    addr, guard, op, rest = int(m.group(1), 16), (m.group(2) or "").strip(), m.group(3), m.group(4)
    base = op.split(".")[0]
    ops = [o.strip() for o in rest.split(",")] if rest else []
    regs = lambda s: re.findall(r"\b(UR\d+|UP\d|R\d+|P\d)\b", s)
    dst, src = [], []
    if guard:
        src += regs(guard)
    nd = 0 if base in ("STS", "BRA", "BSSY", "BSYNC", "ATOMG", "STG") else 1
    if base in ("ISETP", "PLOP3"):
        nd = 2
    if base == "LOP3" and ops and re.match(r"U?P\d", ops[0]):
        nd = 2
    for k, o in enumerate(ops):
        (dst if k < nd else src).extend(r for r in regs(o) if r not in ("PT", "RZ", "URZ", "UPT"))
    wide = 4 if ".128" in op else 2 if ".64" in op else 1
    if base in ("LDS", "LDG") and dst:
        r0 = int(dst[0][1:])
        dst = ["R%d" % (r0 + i) for i in range(wide)]
    if base == "STS" and len(ops) > 1:                       # the stored registers (vector) are inputs
        m2 = re.search(r"R(\d+)", ops[1])
        if m2:
            src += ["R%d" % (int(m2.group(1)) + i) for i in range(wide)]
    return addr, base, dst, src, line.strip()[:90]


def main():
    path, kern, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
    skips = [tuple(int(x, 16) for x in a.split("-")) for a in sys.argv[5:]]
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    body, on = [], False
    for line in sass.splitlines():
        if "Function :" in line:
            on = kern in line
        elif on:
            p = parse(line)
            if p and lo <= p[0] <= hi and not any(a <= p[0] <= b for a, b in skips):
                body.append(p)
    ready, t, marks, crit, prev = {}, 0, [], {}, None
    for it in range(3):
        for addr, base, dst, src, text in body:
            start, why, how = t + 1, prev, "issue order"
            for r in src:
                if r in ready and ready[r][0] > start:
                    start, why, how = ready[r][0], ready[r][1], "waits for " + r
            t = start
            for r in dst:
                ready[r] = (t + LAT.get(base, 5), (addr, it))
            crit[(addr, it)] = (t, why, text, how)
            prev = (addr, it)
        marks.append(t)
    print("instructions per trip: %d   steady-state clocks per trip (one warp, in order): %d" % (len(body), marks[2] - marks[1]))
    # walk the critical chain back from the last instruction of the last copy
    node, chain = (body[-1][0], 2), []
    while node and node[1] >= 1 and len(chain) < 200:
        tt, why, text, how = crit[node]
        chain.append((tt, node[1], text, how))
        node = why
    print("critical chain of the last trip (issue clock, instruction, what it waited for), last instruction first;")
    print("runs of back-to-back `issue order` entries are collapsed:")
    last_how = None
    for tt, it, text, how in chain:
        if how == "issue order" and last_how == "issue order":
            continue
        print("  %5d  #%d  %-62s %s" % (tt, it, text[:62], how))
        last_how = how


if __name__ == "__main__":
    main()
