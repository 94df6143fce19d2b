"""python tools/sass_loops.py FILE.sass: backward branches (loops) of a cuobjdump -sass
listing with the instruction mix of each loop body (dev helper)."""
import re, sys
lines = open(sys.argv[1]).read().splitlines()
ins = []
for l in lines:
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_index = {a: i for i, (a, _) in enumerate(ins)}
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr_index:
            body = ins[addr_index[tgt]: i + 1]
            if len(body) < 40:
                continue
            cnt = {}
            for _, b in body:
                op = b.split()[0] if not b.startswith("@") else b.split()[1]
                op = op.split(".")[0]
                cnt[op] = cnt.get(op, 0) + 1
            fp64 = sum(v for k, v in cnt.items() if k in ("DFMA", "DMUL", "DADD"))
            top = sorted(cnt.items(), key=lambda kv: -kv[1])[:12]
            print("loop %#x..%#x: %d instr, %d FP64; %s" % (tgt, a, len(body), fp64, top))
