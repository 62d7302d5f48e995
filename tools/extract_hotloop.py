"""Extracts the hot loop of a kernel from `cuobjdump -sass`: the innermost backward-branch loop with the most packed-FP32 instructions.
usage: python tools/extract_hotloop.py <lib.so> <mangled-name-substring> [--upto OPCODE] > profiles/xxx.sass
--upto OPCODE: for loops whose body ends in a cold section with loops of its own (k3_rdf_cellw: the candidate pushes and the inlined
drain after the vote), take the densest loop regardless of inner loops and print it up to the first OPCODE."""
import re
import subprocess
import sys

lib, key = sys.argv[1], sys.argv[2]
upto = sys.argv[sys.argv.index("--upto") + 1] if "--upto" in sys.argv else None
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = txt.split("\t\tFunction : ")
body = next(f for f in funcs if f.split("\n", 1)[0].find(key) >= 0)
name = body.split("\n", 1)[0]
ins = []
for ln in body.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_index = {a: i for i, (a, _) in enumerate(ins)}
best = None
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA\S*\s+(?:P\d,\s*)?(?:!?U?P\d+,\s*)?`?\(?0x([0-9a-f]+)", t)
    if m and "BRA" in t:
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in addr_index:
            j = addr_index[tgt]
            loop = ins[j:i + 1]
            score = sum(("FFMA2" in x) or ("FMUL2" in x) for _, x in loop)
            inner = not any(("BRA" in x and k != len(loop) - 1 and re.search(r"0x([0-9a-f]+)", x) and int(re.search(r"0x([0-9a-f]+)", x).group(1), 16) < loop[k][0] and int(re.search(r"0x([0-9a-f]+)", x).group(1), 16) >= tgt) for k, (_, x) in enumerate(loop))
            if score and (inner or (upto and len(loop) < 1500)) and (best is None or score / len(loop) > best[0]):
                best = (score / len(loop), loop)
loop = best[1]
if upto:
    cut = next((k for k, (_, x) in enumerate(loop) if upto in x), len(loop) - 1)
    loop = loop[:cut + 1]
ops = {}
for _, t in loop:
    op = t.split()[1] if t.startswith("@") else t.split()[0]
    op = op.split(".")[0]
    ops[op] = ops.get(op, 0) + 1
print(f"// {name}")
print(f"// hot loop: {len(loop)} instructions per iteration (the loop is unrolled: see the ATOMS count for pairs per iteration)")
print("// opcode mix: " + ", ".join(f"{k} {v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])))
for a, t in loop:
    print(f"/*{a:05x}*/  {t}")
