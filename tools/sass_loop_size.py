#!/usr/bin/env python3
"""Offline proxy for the hot loop's cost: SASS instruction count of the streaming loop of
step_fast_kernel<false,false,1,true> (from the loop head to its backward branch), no GPU needed."""
import re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "shippingenv_b200/libshipenv.so"
ept = sys.argv[2] if len(sys.argv) > 2 else "2"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, lines = None, []
for l in out.splitlines():
    if "Function :" in l:
        fn = l
        continue
    if fn and f"step_fast_kernelILb0ELb0ELi{ept}ELb1" in fn:
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            lines.append((int(m.group(1), 16), m.group(2).strip()))
addr_to_idx = {a: k for k, (a, t) in enumerate(lines)}
for wait in [k for k, (a, t) in enumerate(lines) if "TRYWAIT" in t]:
    found = False
    for k in range(wait + 1, min(wait + 2000, len(lines))):
        a, t = lines[k]
        m = re.search(r"BRA(?:\.U)?\s+(?:\w+,\s*)?0x([0-9a-f]+)", t)
        if m and "DIV" not in t and int(m.group(1), 16) <= lines[wait][0] and int(m.group(1), 16) in addr_to_idx:
            head = addr_to_idx[int(m.group(1), 16)]
            body = lines[head:k + 1]
            if sum("IMAD.WIDE.U32" in t2 for _, t2 in body) < 20:      # the streaming loop contains Philox
                break
            ops = {}
            for _, t2 in body:
                op = t2.split()[1] if t2.startswith("@") else t2.split()[0]
                ops[op.split(".")[0]] = ops.get(op.split(".")[0], 0) + 1
            print(f"streaming loop: {k - head + 1} SASS instructions ({lines[head][0]:#x}..{a:#x})")
            print("  " + ", ".join(f"{o} {c}" for o, c in sorted(ops.items(), key=lambda kv: -kv[1])[:14]))
            alu = sum(c for o, c in ops.items() if o in ("LOP3", "ISETP", "SHF", "PLOP3", "VIADD", "IADD3", "LEA", "SEL", "FSEL",
                                                         "PRMT", "VIMNMX", "MOV", "IABS", "FSETP", "POPC", "VOTE", "SGXT", "BMSK"))
            wide = sum("IMAD.WIDE" in t2 or "IMAD.HI" in t2 for _, t2 in body)
            fma = ops.get("IMAD", 0) - wide
            print(f"  ALU-pipe ops {alu} (x2 cycles), IMAD {fma} (x2), IMAD.WIDE/HI {wide} (x4): "
                  f"ALU {2 * alu} / FMA {2 * fma + 4 * wide} / issue {k - head + 1} cycles per tile of a warp")
            found = True
            break
    if found:
        break
