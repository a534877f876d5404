"""Regenerates profiles/r02_score_loop_sass.txt: the inner loop of pr::score_kernel<8, FMA, 16> from the built object.

    python profiles/score_sass.py  (after python -m dialog_b200.build)
"""
import collections
import re
import subprocess

OBJ = "dialog_b200/_obj/pr_kernels.o"
sass = subprocess.run(["cuobjdump", "-sass", OBJ], capture_output=True, text=True).stdout
m = re.search(r"Function : (\S*score_kernelILi8ELi1ELi16\S*)(.*?)(?=Function : |\Z)", sass, re.S)
lines = []
for ln in m.group(2).splitlines():
    mm = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?;)", ln)
    if mm:
        lines.append((int(mm.group(1), 16), mm.group(2).strip()))
# the loop: the backward branch whose body holds the most FFMA2
best = None
for i, (addr, txt) in enumerate(lines):
    mm = re.search(r"BRA\S*\s+(?:\S+,\s+)?`?\(?(0x[0-9a-f]+)", txt)
    if not mm:
        continue
    tgt = int(mm.group(1), 16)
    if tgt >= addr:
        continue
    body = [t for a, t in lines if tgt <= a <= addr]
    n = sum(1 for t in body if re.match(r"(@\S+\s+)?FFMA2", t))
    if best is None or n > best[0]:
        best = (n, tgt, addr, body)
n, tgt, addr, body = best
hist = collections.Counter(re.sub(r"^@\S+\s+", "", t).split()[0].split(".")[0] for t in body)
ph = 16 * 4 * 8
with open("profiles/r02_score_loop_sass.txt", "w") as f:
    f.write(f"# SASS of the inner loop of pr::score_kernel<H = 8, DOT = FMA, unroll 16> (cuobjdump -sass {OBJ}, CUDA 12.9, sm_100a; profiles/score_sass.py)\n")
    f.write("# One pass of the loop body = 16 steps x 4 points x 8 hypotheses per lane = 512 point-hypotheses per lane.\n")
    f.write(f"# Loop body: {len(body)} instructions, address {tgt:x} .. {addr:x} (back edge: {body[-1]})\n# Opcode histogram of the body:\n")
    for op, k in hist.most_common():
        f.write(f"#   {k:5d}  {op}\n")
    other = len(body) - hist["FFMA2"] - hist["FSET"] - hist["IADD3"]
    f.write(f"# => per point-hypothesis: {hist['FFMA2'] / ph:.3f} FFMA2 (2 FMA each: 3 FMA), {hist['FSET'] / ph:.3f} FSET.BF, {hist['IADD3'] / ph:.3f} IADD3, "
            f"{hist['LDS'] / ph:.4f} LDS.128; {100 * other / len(body):.1f} % of the instructions are neither FMA, compare nor count\n")
    f.write("# No FADD / MOV for |r|: the absolute value is FSET's operand modifier (|R|); the threshold sits in a uniform register.\n#\n")
    f.write("# first 48 instructions of the body\n")
    for a, t in [(a, t) for a, t in lines if tgt <= a <= addr][:48]:
        f.write(f"{a:04x}  {t}\n")
    f.write("# ...\n# last 12 instructions of the body\n")
    for a, t in [(a, t) for a, t in lines if tgt <= a <= addr][-12:]:
        f.write(f"{a:04x}  {t}\n")
print(len(body), dict(hist))
